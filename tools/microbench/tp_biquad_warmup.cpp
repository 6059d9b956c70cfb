// How long a warm-up makes the segments of a time-parallel biquad MERGE with the serial f32 trajectory bit for bit, per cut-off / q /
// input (W = 60 / (1 - pole radius)).  g++ -O2 -ffp-contract=off tp_biquad_warmup.cpp
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
struct C { float b0,b1,b2,a1,a2; };
C lp(uint32_t freq, float q, uint32_t fs, bool hp){ float w0=((2.0f*3.14159265358979323846f)*(float)freq)/(float)fs; float s=sinf(w0), c=cosf(w0); float alpha=s/(2.0f*q);
 float b1,b0,b2; if(!hp){ b1=1.0f-c; b0=b1/2.0f; b2=b0;} else { b1=-1.0f-c; b0=(1.0f+c)/2.0f; b2=b0; }
 float a0=1.0f+alpha, a1=-2.0f*c, a2=1.0f-alpha; return {b0/a0,b1/a0,b2/a0,a1/a0,a2/a0}; }
template<class T> void run(const C&k,const std::vector<float>&x,size_t lo,size_t hi,std::vector<T>&y,size_t st){ T x1=0,x2=0,y1=0,y2=0; for(size_t n=lo;n<hi;n++){ T xn=x[n]; T v=((((T)k.b0*xn)+((T)k.b1*x1))+((T)k.b2*x2))-((T)k.a1*y1); v=v-((T)k.a2*y2); y2=y1;x2=x1;y1=v;x1=xn; if(n>=st) y[n]=v; } }
double radius(const C&k){ double a1=k.a1,a2=k.a2; double disc=a1*a1-4*a2; if(disc<0) return sqrt(a2); double r1=(-a1+sqrt(disc))/2,r2=(-a1-sqrt(disc))/2; return std::max(fabs(r1),fabs(r2)); }
int main(){ const size_t N=96000; std::mt19937_64 g(1); std::uniform_real_distribution<float> u(-1,1);
 struct F{uint32_t f; float q; bool hp;}; 
 for(F fq: {F{600,0.5f,false},F{700,0.5f,false},F{1000,0.5f,false},F{1000,0.707f,false},F{1000,2.0f,false},F{2000,5.0f,false},F{3000,0.5f,false},F{1000,0.5f,true},F{300,0.5f,true},F{8000,0.707f,true}}) for(int sine=0;sine<2;sine++){
  C k=lp(fq.f,fq.q,48000,fq.hp); double r=radius(k); 
  for(double cw: {30.0,60.0,120.0}){ int W=(int)ceil(cw/(1-r)); W=(W+7)/8*8; double worst=0,wref=0; long bad=0,segs=0;
  for(int s=0;s<64;s++){ std::vector<float> x(N); if(sine){ double w=0.001+0.01*fabs(u(g)); for(size_t n=0;n<N;n++) x[n]=(float)sin(w*n)*0.8f; } else for(auto&v:x)v=u(g);
   std::vector<float> ye(N),yt(N); std::vector<double> yd(N); run<float>(k,x,0,N,ye,0); run<double>(k,x,0,N,yd,0);
   const size_t L=6000; for(size_t lo=0;lo<N;lo+=L){ size_t b=lo>=(size_t)W?lo-W:0; run<float>(k,x,b,std::min(N,lo+L),yt,lo); segs++; bool d=false; for(size_t n=lo;n<std::min(N,lo+L);n++) d|= yt[n]!=ye[n]; bad+=d; }
   double peak=0,e1=0,e2=0; for(size_t n=0;n<N;n++){ peak=std::max(peak,std::fabs((double)ye[n])); e1=std::max(e1,std::fabs((double)yt[n]-ye[n])); e2=std::max(e2,std::fabs(yd[n]-ye[n])); }
   worst=std::max(worst,e1/peak); wref=std::max(wref,e2/peak);} 
  printf("%s f=%5u q=%.3f %s r=%.4f W=%5d: segments not bit-exact %4ld/%ld  max|tp-ref|/peak %.2e  (ref vs f64 %.2e)\n",fq.hp?"hp":"lp",fq.f,fq.q,sine?"sine ":"noise",r,W,bad,segs,worst,wref);} } }
