// Fuzzers for the host-side gzip code (no GPU): mutated gzip members through GzStream, random line
// structures through the FASTQ-aware encoder with its decode-and-compare check, under ASan/UBSan.
//   g++ -O1 -g -std=c++17 -fsanitize=address,undefined -fno-sanitize-recover=undefined \
//       -Itrio_binning_b200/csrc -DFUZZ_DECODER -o /tmp/fuzz_dec tools/fuzz_gzip.cpp trio_binning_b200/csrc/gz_stream.cpp trio_binning_b200/csrc/fastq_deflate.cpp -lz && /tmp/fuzz_dec 30000
//   (same with -DFUZZ_ENCODER -o /tmp/fuzz_enc ... && /tmp/fuzz_enc 400)
// Round 1: 30000 mutated inputs (358 still valid, 29642 rejected) and 400 texts (89 encoded and verified,
// 311 declined), no sanitizer report.
#if defined(FUZZ_DECODER)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include <zlib.h>
#include "gz_stream.h"
#include "fastq_deflate.h"
static uint64_t rs=88172645463325252ull; static uint64_t rnd(){ rs^=rs<<13; rs^=rs>>7; rs^=rs<<17; return rs; }
int main(int argc,char**argv){
  int iters = argc>1?atoi(argv[1]):2000;
  // seeds: several kinds of text compressed with zlib at various levels/strategies, plus our own encoder
  std::vector<std::string> texts;
  { std::string t; for(int i=0;i<200;i++){ t+="@r"+std::to_string(i)+"\n"; for(int j=0;j<700;j++) t+="ACGT"[rnd()&3]; t+="\n+\n"; for(int j=0;j<700;j++) t+=(char)(33+rnd()%41); t+="\n"; } texts.push_back(t); }
  { std::string t; for(int i=0;i<300000;i++) t+=(char)((i/1000)%7+'a'); texts.push_back(t); }
  { std::string t; for(int i=0;i<100000;i++) t+=(char)rnd(); texts.push_back(t); }
  std::vector<std::string> seeds;
  for(auto&t:texts) for(int lvl: {0,1,6,9}) for(int strat: {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY}) {
     z_stream zs; memset(&zs,0,sizeof zs); deflateInit2(&zs,lvl,Z_DEFLATED,31,8,strat); std::string o(deflateBound(&zs,t.size())+64,0);
     zs.next_in=(Bytef*)t.data(); zs.avail_in=t.size(); zs.next_out=(Bytef*)&o[0]; zs.avail_out=o.size(); deflate(&zs,Z_FINISH); o.resize(zs.total_out); deflateEnd(&zs); seeds.push_back(o); }
  { std::string o; if (tb::fastq_gzip_member(texts[0].data(), texts[0].size(), o, true)) seeds.push_back(o); }
  long ok=0, err=0;
  std::vector<char> buf(1<<16);
  for(int it=0; it<iters; it++){
     std::string s = seeds[rnd()%seeds.size()];
     int kind = rnd()%4;
     if(kind==0){ int n=1+rnd()%4; for(int i=0;i<n;i++) s[rnd()%s.size()] ^= (char)(1<<(rnd()%8)); }
     else if(kind==1){ s.resize(rnd()%s.size()); }
     else if(kind==2){ size_t p=rnd()%s.size(); size_t l=1+rnd()%64; for(size_t i=p;i<p+l && i<s.size();i++) s[i]=(char)rnd(); }
     else { s += seeds[rnd()%seeds.size()].substr(0, rnd()%2000); }
     tb::GzStream g; g.open_memory((const unsigned char*)s.data(), s.size());
     size_t tot=0; long r;
     while((r=g.read(buf.data(), 1+rnd()%buf.size()))>0) tot+=r;
     if(r<0) err++; else ok++;
  }
  printf("fuzz done: %ld decoded, %ld rejected\n", ok, err);
}
#elif defined(FUZZ_ENCODER)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include "gz_stream.h"
#include "fastq_deflate.h"
static uint64_t rs=0x9E3779B97F4A7C15ull; static uint64_t rnd(){ rs^=rs<<13; rs^=rs>>7; rs^=rs<<17; return rs; }
int main(int argc,char**argv){
  int iters = argc>1?atoi(argv[1]):300; long acc=0, dec=0;
  const char* alph[] = {"ACGT","ACGTN","!\"#$%&'()*+,-./0123456789:;<=>?@ABCDEFGHIJ","~","acgtn","\n","\r\n", "\x01\x02\xff\x80"};
  for(int it=0; it<iters; it++){
    std::string t; int lines = 1 + rnd()%12;
    for(int l=0;l<lines;l++){ const char*a=alph[rnd()%8]; size_t al=strlen(a); size_t len = (rnd()%5==0)? rnd()%700000 : rnd()%30000; int runny = rnd()%3==0;
       char c=a[rnd()%al]; for(size_t i=0;i<len;i++){ if(!runny || rnd()%40==0) c=a[rnd()%al]; t+=c; } t+='\n'; }
    std::string out="prefix"; bool ok = tb::fastq_gzip_member(t.data(), t.size(), out, true);
    if(ok){ acc++; tb::GzStream g; g.open_memory((const unsigned char*)out.data()+6, out.size()-6); std::vector<char> b(t.size()+16); size_t tot=0; long r; while((r=g.read(b.data()+tot, b.size()-tot))>0) tot+=r; if(r<0||tot!=t.size()||memcmp(b.data(),t.data(),tot)){ printf("MISMATCH at iter %d\n", it); return 1; } if (memcmp(out.data(),"prefix",6)) { printf("prefix clobbered\n"); return 1; } }
    else { dec++; if(out!="prefix"){ printf("out modified on decline\n"); return 1; } }
  }
  printf("encoder fuzz: %ld accepted and verified, %ld declined\n", acc, dec);
}
#else
#error define FUZZ_DECODER or FUZZ_ENCODER
#endif
