// test_math.cpp -- TEST INFRASTRUCTURE: pins csrc/mce_math.h (cdiv, cmul, hypot restatements) against the host libgcc / glibc bit for bit.
#include "mce_math.h"
#include <complex.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
typedef double __complex__ C;
static double rnd(int mode){ double u=(rand()+1.0)/(RAND_MAX+2.0); double s=(rand()&1)?1:-1;
  switch(mode){case 0: return s*u; case 1: return s*exp((u-0.5)*80); case 2: return s*exp((u-0.5)*1400); case 3: return (rand()%4==0)?0.0:s*u*1e-12; default: return s*ldexp(u, (rand()%2100)-1074);} }
int main(){ long bad_div=0,bad_mul=0,bad_abs=0,bad_abs2=0,N=4000000; srand(1);
 for(long i=0;i<N;i++){ int mode=i%5; double a=rnd(mode),b=rnd(mode),c=rnd(mode),d=rnd(mode);
  if(i%97==0) d=0; if(i%89==0) c=0; if(i%83==0) a=0; if(i%79==0) b=0;
  C u,v; __real__ u=a; __imag__ u=b; __real__ v=c; __imag__ v=d;
  C q=u/v, p=u*v; double h=cabs(u);
  mce::cplx mq=mce::cdiv(mce::make_cplx(a,b),mce::make_cplx(c,d)), mp=mce::cmul(mce::make_cplx(a,b),mce::make_cplx(c,d));
  double mh=mce::mce_hypot(a,b);
  double qr=__real__ q, qi=__imag__ q, pr=__real__ p, pi=__imag__ p;
  if(memcmp(&qr,&mq.re,8)||memcmp(&qi,&mq.im,8)){ if(!(qr!=qr&&mq.re!=mq.re&&qi!=qi&&mq.im!=mq.im)){ if(bad_div<5) printf("div %a %a %a %a: %a %a vs %a %a\n",a,b,c,d,qr,qi,mq.re,mq.im); bad_div++; } }
  if(memcmp(&pr,&mp.re,8)||memcmp(&pi,&mp.im,8)){ if(!(pr!=pr&&mp.re!=mp.re)) bad_mul++; }
  if(memcmp(&h,&mh,8)){ if(bad_abs<5) printf("abs %a %a: %a vs %a\n",a,b,h,mh); bad_abs++; }
 }
 printf("N=%ld bad_div=%ld bad_mul=%ld bad_abs=%ld\n",N,bad_div,bad_mul,bad_abs); return 0; }
