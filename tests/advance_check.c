// CPU check of the closed form behind advance_bs / fill_run (stringtie_b200/csrc/cov_kernels.cu): m repeated float
// additions bs = fl(bs + c) computed in O(binades crossed).  usage: advance_check [cases]   (exit 1 on any mismatch)
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
static inline uint32_t bits(float x){uint32_t u;memcpy(&u,&x,4);return u;}
static inline float fl(uint32_t u){float x;memcpy(&x,&u,4);return x;}
static float naive(float bs,float c,uint32_t m){volatile float b=bs;for(uint32_t i=0;i<m;i++){b=b+c;}return b;}
static float advance(float bs,float c,uint32_t m){
  while(m){
    volatile float y1=bs+c; m--; if(!m) return y1;
    volatile float y2=y1+c; m--; if(!m) return y2;
    uint32_t b1=bits(y1),b2=bits(y2);
    if((b1^b2)&0xff800000u){bs=y2;continue;}
    volatile float y3=y2+c; m--;
    uint32_t b3=bits(y3);
    if((b3^b2)&0xff800000u){bs=y3;continue;}
    uint32_t inc=b3-b2;
    if(inc==0) return y3;
    uint32_t B=(b3&0xff800000u)+0x00800000u;
    uint32_t t=(B-b3)/inc; if(t>m)t=m;
    bs=fl(b3+t*inc); m-=t;
  }
  return bs;
}
int main(int argc,char**argv){
  long cases=argc>1?atol(argv[1]):300000;
  srand(12345);
  long bad=0,n=0;
  for(long it=0;it<cases;it++){
    float bs,c;
    int mode=rand()%6;
    uint32_t rb=((uint32_t)rand()<<16)^rand(), rc=((uint32_t)rand()<<16)^rand();
    if(mode==0){bs=0.f;} else if(mode==1){bs=(float)(rand()%100000);} else if(mode==2){bs=fl((rb&0x007fffff)|((uint32_t)(100+rand()%60)<<23));}
    else if(mode==3){bs=fl(rb&0x00ffffff);} /* denormal/small */ else bs=(float)(rand()%1000)/3.0f;
    int cm=rand()%6;
    if(cm==0) c=1.0f/3.0f; else if(cm==1) c=(float)(rand()%50)/7.0f; else if(cm==2) c=fl((rc&0x007fffff)|((uint32_t)(90+rand()%60)<<23));
    else if(cm==3) c=fl(rc&0x00ffffff); else if(cm==4) c=2.98e-8f*(rand()%5); else c=(float)(rand()%64)*0.25f+0.5f*(rand()%2);
    /* tie cases: c = exactly half ulp of bs multiples */
    if(it%7==0 && bs>0){ uint32_t e=(bits(bs)>>23)&0xff; if(e>30){ c=fl(((e-24)<<23)) * (float)(1+2*(rand()%9)); } }
    uint32_t m=1+rand()%10500;
    float a=naive(bs,c,m), b=advance(bs,c,m);
    n++;
    if(bits(a)!=bits(b)){ if(bad<10) printf("MISMATCH bs=%a c=%a m=%u naive=%a adv=%a\n",bs,c,m,a,b); bad++; }
  }
  printf("%ld cases, %ld mismatches\n",n,bad);
  return bad!=0;
}
