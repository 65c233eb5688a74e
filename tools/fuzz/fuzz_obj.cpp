// Mutates the OBJ files a fixed scene refers to (ASan + UBSan build).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <random>
#include <sys/stat.h>
#include "svgf_b200.h"
static std::string slurp(const char *p){ FILE*f=fopen(p,"rb"); std::string v; if(!f) return v; fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); v.resize(n); if(fread(&v[0],1,n,f)!=(size_t)n) v.clear(); fclose(f); return v; }
int main(int argc,char**argv){
  int iters=atoi(argv[1]); unsigned seed=atoi(argv[2]); const char*scene=argv[3]; const char*models=argv[4];
  std::mt19937 rng(seed); int ok=0,bad=0;
  static const char *words[]={"v ","vn ","vt ","f ","1/2/3 ","-1//-1 ","9999999 ","0 ","/","//","1e38 ","nan ","inf ","-1e38 ","1e-45 ","\n","g x\n","usemtl m\n"," "};
  char dir[64]; snprintf(dir,sizeof dir,"/tmp/svgf_fuzz_m_%u",seed); mkdir(dir,0755);
  std::vector<std::string> names, bases;
  for(int a=5;a<argc;a++){ names.push_back(argv[a]); bases.push_back(slurp((std::string(models)+"/"+argv[a]).c_str())); if(bases.back().empty()){printf("cannot read %s\n",argv[a]);return 1;} }
  for(int it=0;it<iters;it++){
    for(size_t a=0;a<names.size();a++){
      std::string v=bases[a]; int nmut=(rng()%4==0)?0:1+rng()%6;
      for(int k=0;k<nmut;k++){
        int mode=rng()%5; size_t pos=v.empty()?0:rng()%v.size();
        if(mode==0&&!v.empty()) v.resize(pos);
        else if(mode==1&&!v.empty()) v[pos]=(char)(32+rng()%95);
        else if(mode==2){ v.insert(pos,words[rng()%(sizeof words/sizeof*words)]); }
        else if(mode==3&&!v.empty()){ size_t n=rng()%40; if(pos+n<v.size()) v.erase(pos,n); }
        else if(!v.empty()) v[pos]=(char)rng();
      }
      FILE*f=fopen((std::string(dir)+"/"+names[a]).c_str(),"wb"); fwrite(v.data(),1,v.size(),f); fclose(f);
    }
    svgf_scene*s=nullptr; int rc=svgf_scene_load(&s,scene,dir);
    if(rc==0&&s){ svgf_scene_desc d; unsigned char px[48]={0}; for(int i=0;i<svgf_scene_num_textures(s);i++) svgf_scene_set_texture(s,i,4,4,3,px); int r2=svgf_scene_describe(s,64,64,&d); if(r2==0) ok++; else bad++; float b[600]; svgf_scene_mesh_boxes(s,b,100);} else bad++;
    if(s) svgf_scene_free(s);
  }
  printf("loaded %d rejected %d\n",ok,bad); return 0;
}
