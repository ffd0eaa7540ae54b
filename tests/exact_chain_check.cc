// CPU check of the integer form of a float addition chain (stringtie_b200/csrc/exact_chain.cuh): blocks of 32 x E terms
// accepted by the same increment function and the same bounds as warp_chain_add must give the bits of the plain loop.
// usage: exact_chain_check [cases]   (exit 1 on any mismatch; prints the share of blocks that took the integer form)
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include "../stringtie_b200/csrc/exact_chain.cuh"
static inline uint32_t bits(float x){uint32_t u;memcpy(&u,&x,4);return u;}
static inline float fl(uint32_t u){float x;memcpy(&x,&u,4);return x;}
static float plain(float s,const float*x,int n){volatile float a=s;for(int i=0;i<n;i++){a=a+x[i];}return a;}
static long fast_blocks=0, all_blocks=0, tie_terms=0;
// one block the way the warp does it: lanes own E consecutive terms and run them for both parities of the sum in front,
// the lanes' maps are composed in order (the warp scan does the same with a tree)
static float block(float s,const float*x,int E){
  all_blocks++;
  const uint32_t sb=bits(s); const int ex_s=(int)((sb>>23)&0xffu);
  if(ex_s>=1&&ex_s<=254){
    const uint32_t neg_s=sb>>31; bool bad=false;
    EcLane L[32];
    for(int l=0;l<32;l++) for(int e=0;e<E;e++){ bool tie; const int32_t r=ec_chain_incr(bits(x[l*E+e]),ex_s,neg_s,tie,bad); ec_lane_step(L[l],e==0,r,tie); }
    /* inclusive maps by the same doubling steps as the warp scan */
    int32_t i0[32],i1[32];
    for(int l=0;l<32;l++){i0[l]=L[l].run;i1[l]=L[l].run+L[l].delta;}
    for(int o=1;o<32;o<<=1){ int32_t n0[32],n1[32]; for(int l=0;l<32;l++){ if(l>=o) ec_compose(i0[l-o],i1[l-o],i0[l],i1[l],n0[l],n1[l]); else {n0[l]=i0[l];n1[l]=i1[l];} } memcpy(i0,n0,sizeof n0); memcpy(i1,n1,sizeof n1); }
    const int32_t A=(int32_t)((sb&0x7fffffu)|0x800000u);
    int32_t fin=0;
    for(int l=0;l<32;l++){
      const int32_t at=A+(l==0?0:((A&1)?i1[l-1]:i0[l-1]));
      if(at+L[l].mn<EC_LO||at+L[l].mx>EC_HI) bad=true;
      if(l==31) fin=at+L[l].run+((at&1)?L[l].delta:0);
    }
    if(!bad){ fast_blocks++; for(int i=0;i<32*E;i++){bool t,bb=false; ec_chain_incr(bits(x[i]),ex_s,neg_s,t,bb); tie_terms+=t;} return fl((sb&0xff800000u)|((uint32_t)fin&0x7fffffu)); }
  }
  return plain(s,x,32*E);
}
static uint32_t rnd(){return ((uint32_t)rand()<<16)^(uint32_t)rand();}
int main(int argc,char**argv){
  long cases=argc>1?atol(argv[1]):20000;
  srand(4242);
  long badn=0;
  static float x[32*8*64];
  const float wts[8]={1.0f,0.5f,1.0f/3.0f,0.25f,0.2f,1.0f/7.0f,0.1f,2.0f};
  for(long it=0;it<cases;it++){
    const int E=(it&1)?8:4, nb=1+rand()%64, n=nb*32*E;
    const int mode=rand()%9;
    float s;
    const int sm=rand()%7;
    if(sm==0) s=0.f; else if(sm==1) s=(float)(rand()%3000000); else if(sm==2) s=fl((rnd()&0x007fffff)|((uint32_t)(100+rand()%60)<<23));
    else if(sm==3) s=-(float)(rand()%100000)/3.0f; else if(sm==4) s=fl(((uint32_t)(120+rand()%30))<<23); /* power of two */
    else if(sm==5) s=16777216.0f-(float)(rand()%4000); else s=(float)(rand()%1000)/7.0f;
    if(mode==8) s=(float)(16777216+rand()%100000000)*(float)(1+rand()%16)*((rand()&1)?1.f:-1.f);
    for(int i=0;i<n;i++){
      float v;
      switch(mode){
        case 0: v=wts[rand()%8]; break;                                        /* coverage deposits, positive */
        case 1: v=wts[rand()%8]*((rand()&1)?1.f:-1.f); break;                  /* +w / -w mixed */
        case 2: v=wts[rand()%5]*(float)(1+rand()%300)*((float)(rand()%1000)/999.0f); break;   /* rprop * bp * readcov */
        case 3: v=fl((rnd()&0x007fffff)|((uint32_t)(100+rand()%40)<<23)); if(rand()&1) v=-v; break;  /* wild exponents */
        case 4: { uint32_t e=(bits(s)>>23)&0xff; if(e<30) e=127; v=fl((e-24)<<23)*(float)(1+2*(rand()%5)); if(rand()%3==0) v=-v; } break; /* half grid steps: ties */
        case 5: v=(rand()%50==0)?fl(rnd()):wts[rand()%8]; break;               /* NaN / Inf / denormals now and then */
        case 6: v=-wts[rand()%8]; break;                                       /* draining a sum */
        case 7: v=(rand()%20==0)?0.f:1.0f; break;
        default: v=(float)(1+rand()%150)*wts[rand()%4]; break;   /* bases x weight on a sum beyond 2^24: ties all the time */
      }
      x[i]=v;
    }
    float a=s,b=s;
    for(int k=0;k<nb;k++){ a=plain(a,x+k*32*E,32*E); b=block(b,x+k*32*E,E); if(bits(a)!=bits(b)&&!(a!=a&&b!=b)){ if(badn<10) printf("MISMATCH case %ld block %d mode %d: plain=%a block=%a\n",it,k,mode,a,b); badn++; break; } }
  }
  printf("%ld cases, %ld mismatches, %ld of %ld blocks in integer form (%ld ties among their terms)\n",cases,badn,fast_blocks,all_blocks,tie_terms);
  return badn!=0;
}
