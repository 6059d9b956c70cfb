// CPU experiment behind DESIGN.md section 4.4 ("why the recurrence is not cut in time"): run the reference's f32
// Direct-Form-I biquad (src/source/blt.rs:558-560 operation order) twice over the same input -- once from the start of
// the stream, once from zero state at a later sample -- and measure after how many samples the two trajectories become
// bit-identical (from then on they stay so).  A speculative time split is only exact if that happens within the warm-up.
//   g++ -O2 -ffp-contract=off -o biquad_merge tools/microbench/biquad_merge.cpp && ./biquad_merge
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
#include <algorithm>
struct Co { float b0,b1,b2,a1,a2; };
static Co lp(unsigned freq, float q, unsigned fs){
  const float PI=3.14159265358979323846f; float w0 = 2.0f*PI*(float)freq/(float)fs; float alpha = sinf(w0)/(2.0f*q); float c = cosf(w0);
  float b1 = 1.0f-c, b0=b1/2.0f, b2=b0; float a0=1.0f+alpha, a1=-2.0f*c, a2=1.0f-alpha; return {b0/a0,b1/a0,b2/a0,a1/a0,a2/a0}; }
static void run(const Co& k, const float* x, size_t n, size_t s, std::vector<float>& y){
  y.assign(n,0.f); float x1=0,x2=0,y1=0,y2=0;
  for(size_t i=s;i<n;i++){ float xn=x[i]; float v = k.b0*xn; v = v + k.b1*x1; v = v + k.b2*x2; v = v - k.a1*y1; v = v - k.a2*y2; x2=x1;x1=xn;y2=y1;y1=v;y[i]=v; } }
int main(){
  unsigned fs=48000; size_t n=48000;
  for(int kind=0;kind<3;kind++) for(unsigned f: {200u,1000u}){
    Co k=lp(f,0.5f,fs); std::vector<size_t> merges; int N=600;
    for(int seed=0;seed<N;seed++){ std::mt19937_64 g(0x5EED+seed); std::uniform_real_distribution<float> U(-1,1);
      std::vector<float> x(n); for(size_t i=0;i<n;i++) x[i]= kind==0? U(g) : kind==1? 0.8f*sinf(6.2831853f*(100.f+seed*7.3f)*i/48000.f) : 0.3f*U(g)+0.5f*sinf(6.2831853f*(50.f+seed)*i/48000.f);
      std::vector<float> yf,ys; run(k,x.data(),n,0,yf); size_t s=4001+seed; run(k,x.data(),n,s,ys);
      size_t last=s; bool any=false; for(size_t i=s;i<n;i++) if(yf[i]!=ys[i]){last=i;any=true;}
      merges.push_back(any? last-s+1:0);
    }
    std::sort(merges.begin(),merges.end());
    auto frac=[&](size_t W){ size_t c=0; for(auto m:merges) if(m>W) c++; return (double)c/N; };
    printf("kind=%d lp(%u): median merge %zu  p90 %zu  p99 %zu | P(unmerged) W=1000: %.3f 1500: %.3f 2000: %.3f 3000: %.3f 5000: %.3f 10000: %.3f 40000: %.3f\n", kind,f, merges[N/2], merges[N*9/10], merges[N*99/100], frac(1000),frac(1500),frac(2000),frac(3000),frac(5000),frac(10000),frac(40000));
  }
}
