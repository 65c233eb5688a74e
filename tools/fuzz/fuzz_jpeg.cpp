// Mutation fuzzer for svgf_jpeg_decode (ASan + UBSan build).
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <random>
bool svgf_jpeg_decode(const unsigned char*, size_t, int*, int*, int*, std::vector<unsigned char>&, std::string&);
static std::vector<unsigned char> slurp(const char *p){ FILE*f=fopen(p,"rb"); std::vector<unsigned char> v; if(!f) return v; fseek(f,0,SEEK_END); long n=ftell(f); fseek(f,0,SEEK_SET); v.resize(n); if(fread(v.data(),1,n,f)!=(size_t)n) v.clear(); fclose(f); return v; }
int main(int argc,char**argv){
  int iters=atoi(argv[1]); unsigned seed=atoi(argv[2]);
  std::mt19937 rng(seed);
  int ok=0, bad=0;
  for(int a=3;a<argc;a++){
    auto base=slurp(argv[a]); if(base.empty()){printf("cannot read %s\n",argv[a]);return 1;}
    for(int it=0;it<iters;it++){
      auto v=base;
      int mode=rng()%6;
      // most damage in the headers (first 1 KB) where the structure lives
      size_t span = (rng()%3)? std::min<size_t>(v.size(), 1200): v.size();
      int nmut = 1+rng()%8;
      if(mode==0){ v.resize(rng()%v.size()); }
      else if(mode==1){ for(int k=0;k<nmut;k++) v[rng()%span]=(unsigned char)rng(); }
      else if(mode==2){ for(int k=0;k<nmut;k++) v[rng()%span]^=(unsigned char)(1u<<(rng()%8)); }
      else if(mode==3){ for(int k=0;k<nmut;k++) v[rng()%span]=(rng()&1)?0xFF:0x00; }
      else if(mode==4){ size_t p=rng()%span, n=rng()%64; if(p+n<v.size()) v.erase(v.begin()+p, v.begin()+p+n); }
      else { size_t p=rng()%span; size_t n=rng()%64; std::vector<unsigned char> ins(n); for(auto&b:ins) b=(unsigned char)rng(); v.insert(v.begin()+p, ins.begin(), ins.end()); }
      int w=0,h=0,c=0; std::vector<unsigned char> out; std::string err;
      bool r=svgf_jpeg_decode(v.data(), v.size(), &w,&h,&c,out,err);
      if(r){ ok++; if(out.size()!=(size_t)w*h*3){printf("size mismatch\n");return 2;} } else bad++;
    }
  }
  printf("decoded %d rejected %d\n",ok,bad);
  return 0;
}
