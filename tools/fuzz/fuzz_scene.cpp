// Mutation fuzzer for svgf_scene_load + svgf_scene_describe (host only), ASan + UBSan build.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <random>
#include "svgf_b200.h"
static std::string slurp(const char *p){ FILE*f=fopen(p,"rb"); std::string v; if(!f) return v; fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); v.resize(n); if(fread(&v[0],1,n,f)!=(size_t)n) v.clear(); fclose(f); return v; }
int main(int argc,char**argv){
  int iters=atoi(argv[1]); unsigned seed=atoi(argv[2]); const char*models=argv[3];
  std::mt19937 rng(seed); int ok=0,bad=0;
  static const char *words[]={"MATERIAL","OBJECT","CAMERA","TEXTURE","RGB","SPECEX","SPECRGB","REFL","REFR","REFRIOR","EMITTANCE","cube","sphere","mesh","material","TRANS","ROTAT","SCALE","RES","FOVY","ITERATIONS","DEPTH","FILE","EYE","LOOKAT","UP","-1","0","1e38","nan","inf","999999999999","-0","1e-45"," ","\n","\r\n","\t"};
  char path[64]; snprintf(path,sizeof path,"/tmp/svgf_fuzz_s_%u.txt",seed);
  for(int a=4;a<argc;a++){
    std::string base=slurp(argv[a]); if(base.empty()){printf("cannot read %s\n",argv[a]);return 1;}
    for(int it=0;it<iters;it++){
      std::string v=base; int nmut=1+rng()%6;
      for(int k=0;k<nmut;k++){
        int mode=rng()%6; size_t pos=v.empty()?0:rng()%v.size();
        if(mode==0&&!v.empty()) v.resize(pos);
        else if(mode==1&&!v.empty()) v[pos]=(char)(32+rng()%95);
        else if(mode==2){ const char*w=words[rng()%(sizeof words/sizeof*words)]; v.insert(pos,w); }
        else if(mode==3&&!v.empty()){ size_t n=rng()%40; if(pos+n<v.size()) v.erase(pos,n); }
        else if(mode==4&&!v.empty()){ // swap a number for an extreme one
          const char*w=words[26+rng()%8]; size_t e=v.find_first_of(" \n",pos); if(e!=std::string::npos) v.replace(pos,e-pos,w); }
        else if(!v.empty()) v[pos]=(char)rng();
      }
      FILE*f=fopen(path,"wb"); fwrite(v.data(),1,v.size(),f); fclose(f);
      svgf_scene*s=nullptr; int rc=svgf_scene_load(&s,path,models);
      if(rc==0&&s){ svgf_scene_desc d; float e[3],l[3],u[3],fov; int res[2]; svgf_scene_camera(s,e,l,u,&fov,res);
        int nt=svgf_scene_num_textures(s); std::vector<unsigned char> px(4*4*3,7); for(int i=0;i<nt;i++){ svgf_scene_texture_file(s,i); svgf_scene_set_texture(s,i,4,4,3,px.data()); }
        int r2=svgf_scene_describe(s,64,64,&d); if(r2==0) ok++; else bad++; float b[600]; svgf_scene_mesh_boxes(s,b,100);} else { bad++; if(s) svgf_scene_error(s);} 
      if(s) svgf_scene_free(s);
    }
  }
  printf("loaded %d rejected %d\n",ok,bad); return 0;
}
