#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
/* CPU check of the fp64 cbrt / expm1 / tanh restatements that xtensor_b200/csrc/xtb_ops.cuh runs on the device
 * (glibc_cbrt / glibc_expm1 / glibc_tanh): the same operation sequences, here against this image's libm.
 *   gcc -O2 -mfma -ffp-contract=off -fno-fast-math -fno-builtin tools/glibc_replica_check.c -o /tmp/chk -lm && /tmp/chk
 * prints the number of results that differ from libm's bit for bit (expected: all zero).  Test infrastructure. */
#define FMA __builtin_fma
static inline uint32_t hi(double x){uint64_t u;memcpy(&u,&x,8);return u>>32;}
static inline uint32_t lo(double x){uint64_t u;memcpy(&u,&x,8);return (uint32_t)u;}
static inline double mk(uint32_t h){uint64_t u=((uint64_t)h)<<32; double d; memcpy(&d,&u,8); return d;}
static double my_cbrt(double x) {
    int xe; double xm = frexp(fabs(x), &xe);
    if (xe == 0 && (x == 0.0 || !(fabs(x) <= 1.7976931348623157e308))) return x + x;
    double u = (0.354895765043919860 + ((1.50819193781584896 - ((2.11499494167371287 - ((2.44693122563534430 - ((1.83469277483613086 - (0.784932344976639262 - 0.145263899385486377 * xm) * xm) * xm)) * xm)) * xm)) * xm));
    double t2 = u * u * u, f = 1.0; int r = xe % 3;
    if (r == -2) f = 1.0 / 1.5874010519681994748; else if (r == -1) f = 1.0 / 1.2599210498948731648; else if (r == 1) f = 1.2599210498948731648; else if (r == 2) f = 1.5874010519681994748;
    double ym = u * (t2 + 2.0 * xm) / (2.0 * t2 + xm) * f;
    return ldexp(x > 0.0 ? ym : -ym, xe / 3);
}
static double my_expm1(double x) {
    const double one=1.0, huge=1.0e+300, tiny=1.0e-300, o_threshold=7.09782712893383973096e+02,
    ln2_hi=6.93147180369123816490e-01, ln2_lo=1.90821492927058770002e-10, invln2=1.44269504088896338700e+00,
    Q1=-3.33333333333331316428e-02, Q2=1.58730158725481460165e-03, Q3=-7.93650757867487942473e-05,
    Q4=4.00821782732936239552e-06, Q5=-2.01099218183624371326e-07;
    double y,hi_,lo_,c=0,t,e,hxs,hfx,r1,twopk; int32_t k; uint32_t hx=hi(x), xsb=hx&0x80000000; hx&=0x7fffffff;
    if(hx>=0x4043687A){ if(hx>=0x40862E42){ if(hx>=0x7ff00000){ if(((hx&0xfffff)|lo(x))!=0) return x+x; else return xsb==0?x:-1.0;} if(x>o_threshold) return huge*huge;} if(xsb!=0){ if(x+tiny<0.0) return tiny-one; } }
    if(hx>0x3fd62e42){ if(hx<0x3FF0A2B2){ if(xsb==0){hi_=x-ln2_hi;lo_=ln2_lo;k=1;} else {hi_=x+ln2_hi;lo_=-ln2_lo;k=-1;} }
        else { k=(int32_t)FMA(invln2,x,(xsb==0)?0.5:-0.5); t=k; hi_=FMA(-t,ln2_hi,x); lo_=t*ln2_lo; } x=hi_-lo_; c=(hi_-x)-lo_; }
    else if(hx<0x3c900000){ t=huge+x; return x-(t-(huge+x)); }
    else k=0;
    hfx=0.5*x; hxs=x*hfx;
    { double R1=FMA(hxs,Q1,one), h2=hxs*hxs, R2=FMA(hxs,Q3,Q2), h4=h2*h2, R3=FMA(hxs,Q5,Q4); r1=FMA(h4,R3,FMA(h2,R2,R1)); }
    t=FMA(-r1,hfx,3.0); e=hxs*((r1-t)/FMA(-x,t,6.0));
    if(k==0) return x-FMA(x,e,-hxs);
    else { twopk=mk(0x3ff00000+(k<<20)); e=FMA(x,(e-c),-c); e-=hxs;
      if(k==-1) return FMA(0.5,(x-e),-0.5);
      if(k==1){ if(x<-0.25) return -2.0*(e-(x+0.5)); else return FMA(2.0,(x-e),one);} 
      if(k<=-2||k>56){ y=one-(e-x); if(k==1024) y=y*2.0*0x1p1023; else y=y*twopk; return y-one; }
      t=one;
      if(k<20){ t=mk(0x3ff00000-(0x200000>>k)); y=t-(e-x); y=y*twopk; }
      else { t=mk((0x3ff-k)<<20); y=x-(e+t); y+=one; y=y*twopk; } }
    return y;
}
static double my_tanh(double x){
    const double one=1.0,two=2.0,tiny=1.0e-300; double t,z; int32_t jx=(int32_t)hi(x), ix=jx&0x7fffffff; uint32_t lx=lo(x);
    if(ix>=0x7ff00000){ if(jx>=0) return one/x+one; else return one/x-one; }
    if(ix<0x40360000){ if((ix|lx)==0) return x; if(ix<0x3c800000) return x*(one+x);
      if(ix>=0x3ff00000){ t=my_expm1(two*fabs(x)); z=one-two/(t+two);} else { t=my_expm1(-two*fabs(x)); z=-t/(t+two);} }
    else z=one-tiny;
    return (jx>=0)?z:-z;
}
int main(){
    srand48(2); long bad_e=0,bad_t=0,bad_e2=0,bad_c=0; const long N=8000000;
    for(long i=0;i<N;i++){
        double xe=(drand48()*2-1)*60; if(my_expm1(xe)!=expm1(xe)) bad_e++;
        double xs=(drand48()*2-1)*ldexp(1.0,(int)(drand48()*70)-60); double a=my_expm1(xs), b=expm1(xs); if(memcmp(&a,&b,8)) bad_e2++;
        double xt=(drand48()*2-1)*25; if(my_tanh(xt)!=tanh(xt)) bad_t++;
        double xc=(drand48()*2-1)*ldexp(1.0,(int)(drand48()*600)-300); { double a=my_cbrt(xc), b=cbrt(xc); if(memcmp(&a,&b,8)) bad_c++; }
    }
    double sp[]={0.0,-0.0,INFINITY,-INFINITY,710.0,-800.0,1e-320,0.34657359027997264,1.0397207708399179,709.78,38.8,-38.8,56*0.6931471805599453};
    int bs=0; for(unsigned i=0;i<sizeof(sp)/8;i++){ double a=my_expm1(sp[i]),b=expm1(sp[i]); if(memcmp(&a,&b,8)) bs++; a=my_tanh(sp[i]); b=tanh(sp[i]); if(memcmp(&a,&b,8)) bs++; }
    printf("mismatches vs libm: cbrt %ld, expm1 wide %ld small-range %ld, tanh %ld (of %ld each), specials %d\n",bad_c,bad_e,bad_e2,bad_t,N,bs);
    return 0;
}
