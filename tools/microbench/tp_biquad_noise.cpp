// Exact f32 biquad (src/source/blt.rs:558-560 order) against segments restarted from zero state W samples early, and against an f64 run:
// which cut-offs a time-parallel plan may serve within 1e-5 * peak.  g++ -O2 -ffp-contract=off tp_biquad_noise.cpp
// time-parallel biquad experiment: exact f32 DF1 (blt.rs order) vs segments restarted from zero state W samples early
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
struct C { float b0,b1,b2,a1,a2; };
C lp(uint32_t freq, float q, uint32_t fs){ float w0=((2.0f*3.14159265358979323846f)*(float)freq)/(float)fs; float s=sinf(w0), c=cosf(w0); float alpha=s/(2.0f*q);
 float b1=1.0f-c, b0=b1/2.0f, b2=b0, a0=1.0f+alpha, a1=-2.0f*c, a2=1.0f-alpha; return {b0/a0,b1/a0,b2/a0,a1/a0,a2/a0}; }
template<class T> void run(const C&k,const std::vector<float>&x,size_t lo,size_t hi,std::vector<T>&y,size_t st){ T x1=0,x2=0,y1=0,y2=0; for(size_t n=lo;n<hi;n++){ T xn=x[n]; T v=((((T)k.b0*xn)+((T)k.b1*x1))+((T)k.b2*x2))-((T)k.a1*y1); v=v-((T)k.a2*y2); y2=y1;x2=x1;y1=v;x1=xn; if(n>=st) y[n]=v; } }
int main(){ const size_t N=96000; std::mt19937_64 g(1); std::uniform_real_distribution<float> u(-1,1);
 for(uint32_t f: {100u,200u,300u,500u,700u,1000u,2000u,5000u}) for(int sine=0;sine<2;sine++){
  double worst_tp=0,worst_ref=0,worst_tp64=0; double a2v=0; 
  for(int s=0;s<24;s++){ std::vector<float> x(N); if(sine){ double w=0.001+0.01*u(g); for(size_t n=0;n<N;n++) x[n]=(float)sin(w*n)*0.8f; } else for(auto&v:x)v=u(g);
   C k=lp(f,0.5f,48000); a2v=k.a2; std::vector<float> ye(N),yt(N); std::vector<double> yd(N); run<float>(k,x,0,N,ye,0); run<double>(k,x,0,N,yd,0);
   for(int W: {4096}){ const size_t L=6000; for(size_t lo=0;lo<N;lo+=L){ size_t st=lo, b=lo>=(size_t)W?lo-W:0; run<float>(k,x,b,std::min(N,lo+L),yt,st);} }
   double peak=0,e1=0,e2=0,e3=0; for(size_t n=0;n<N;n++){ peak=std::max(peak,std::fabs((double)ye[n])); e1=std::max(e1,std::fabs((double)yt[n]-ye[n])); e2=std::max(e2,std::fabs(yd[n]-ye[n])); e3=std::max(e3,std::fabs(yd[n]-yt[n])); }
   worst_tp=std::max(worst_tp,e1/peak); worst_ref=std::max(worst_ref,e2/peak); worst_tp64=std::max(worst_tp64,e3/peak);} 
  printf("f=%5u %s a2=%.5f  |tp-f32ref|/peak %.2e   |f32ref-f64|/peak %.2e   |tp-f64|/peak %.2e\n",f,sine?"sine ":"noise",a2v,worst_tp,worst_ref,worst_tp64);} }
