/* Exactness check of the FP64 error-free modular product used by the engine's FpArith (csrc/ntt.cuh):
 * for q < 2^47, |y| < 2^51:  r = fma(-k, q, p) + e  with p = y*w rounded, e = fma(y, w, -p), k = rint(y * (w/q))
 * is an integer congruent to y*w mod q with |r| < 0.63 q.  Compiled and run by tests/test_fp_modmul.py. */
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <stdlib.h>
typedef unsigned __int128 u128; typedef __int128 i128;
static const double MAGIC = 6755399441055744.0; /* 1.5*2^52 */
static inline double modmul(double y, double w, double winv, double q){
  double k = fma(y, winv, MAGIC) - MAGIC;
  double p = y*w;
  double e = fma(y, w, -p);
  double r = fma(-k, q, p);
  return r + e;
}
int main(){
  uint64_t seed=88172645463325252ULL; long bad=0; double maxr=0;
  for(int bq=30;bq<=47;bq++){
    for(int it=0;it<400000;it++){
      seed^=seed<<13; seed^=seed>>7; seed^=seed<<17;
      uint64_t q=((seed>>3)%(1ULL<<(bq-1)))+(1ULL<<(bq-1)); q|=1;
      seed^=seed<<13; seed^=seed>>7; seed^=seed<<17;
      uint64_t w=seed%q;
      seed^=seed<<13; seed^=seed>>7; seed^=seed<<17;
      int lazy_bits = bq+4; if(lazy_bits>51) lazy_bits=51;
      int64_t y=(int64_t)(seed%(1ULL<<lazy_bits)); if(seed&(1ULL<<63)) y=-y;
      double winv=(double)w/(double)q;
      double r=modmul((double)y,(double)w,winv,(double)q);
      i128 exact=(i128)y*(i128)w; i128 rr=(i128)r; 
      i128 diff=exact-rr; i128 m=diff%(i128)q; 
      if(m!=0 || r!=floor(r)) {bad++; if(bad<5) printf("BAD bq=%d q=%lu w=%lu y=%ld r=%f\n",bq,q,w,y,r);}
      double ar=fabs(r)/(double)q; if(ar>maxr)maxr=ar;
    }
  }
  printf("bad=%ld max|r|/q=%f\n",bad,maxr);
  return 0;
}
