// Exhaustive check of the sinf restatement used by k_siggen (rb_kernels.cu sinf_glibc_0_tau) against the libm of this machine:
// every float of [0, 2*pi_f32], with and without contraction of the double multiply-adds.
// g++ -O2 -ffp-contract=off -mfma -pthread tools/microbench/sinf_exhaustive.cpp -o /tmp/sinf_x && /tmp/sinf_x
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <thread>
#include <vector>
struct T { double sign[4]; double hpi_inv, hpi, c0,c1,c2,c3,c4,s1,s2,s3; };
static const T tab[2] = {
 {{1.0,-1.0,-1.0,1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2, 0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
 {{1.0,-1.0,-1.0,1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2, -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3, 0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
template<bool FMA> static inline double mad(double a,double b,double c){ if(FMA) return std::fma(a,b,c); volatile double p=a*b; return p+c; }
template<bool FMA> static inline float poly(double x,double x2,const T*p,int n){
  if((n&1)==0){ double x3=x*x2; double s1=mad<FMA>(x2,p->s3,p->s2); double x7=x3*x2; double s=mad<FMA>(x3,p->s1,x); return (float)mad<FMA>(x7,s1,s);}
  else { double x4=x2*x2; double c2=mad<FMA>(x2,p->c4,p->c3); double c1=mad<FMA>(x2,p->c1,p->c0); double x6=x4*x2; double c=mad<FMA>(x4,p->c2,c1); return (float)mad<FMA>(x6,c2,c);}
}
static inline uint32_t top12(float f){uint32_t u; memcpy(&u,&f,4); return (u>>20)&0x7ff;}
template<bool FMA> static float my_sinf(float y){
  double x=y; const T*p=&tab[0];
  if(top12(y)<top12(0x1.921FB6p-1f)){ double s=x*x; if(top12(y)<top12(0x1p-12f)) return y; return poly<FMA>(x,s,p,0);}
  double r=x*p->hpi_inv; int n=((int32_t)r+0x800000)>>24; double xr=mad<FMA>(-(double)n,p->hpi,x);
  double s=p->sign[n&3]; if(n&2) p=&tab[1];
  return poly<FMA>(xr*s,xr*xr,p,n);
}
int main(int argc,char**argv){ const uint64_t stride = argc>1 ? strtoull(argv[1],0,10) : 1;
  float tau=6.2831855f; uint32_t hi; memcpy(&hi,&tau,4);
  const int NT=64; std::atomic<uint64_t> bad0{0},bad1{0}; std::vector<std::thread> th;
  uint32_t first0=0,first1=0;
  for(int t=0;t<NT;t++) th.emplace_back([&,t]{ uint64_t b0=0,b1=0; for(uint64_t u=(uint64_t)t*stride;u<=hi;u+=(uint64_t)NT*stride){ float y; uint32_t uu=(uint32_t)u; memcpy(&y,&uu,4); float g=sinf(y); float a=my_sinf<false>(y), b=my_sinf<true>(y); if(memcmp(&g,&a,4)){b0++; first0=uu;} if(memcmp(&g,&b,4)){b1++; first1=uu;} } bad0+=b0; bad1+=b1;});
  for(auto&x:th)x.join();
  printf("checked every %llu-th of %u floats: mismatches no-fma %llu (e.g. %08x), fma %llu (e.g. %08x)\n",(unsigned long long)stride,hi+1,(unsigned long long)bad0.load(),first0,(unsigned long long)bad1.load(),first1);
}
