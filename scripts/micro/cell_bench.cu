// Microbenchmark: cost of the cell-index computation of the voxelizer's scatter kernel.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
struct Prm { double r[3], v[3]; float rf[3], vf[3], inv_vf[3], rv_abs[3]; int g[3]; int regime; };
template <int MODE>
__device__ __forceinline__ bool axis_cell(const Prm &q, int j, float p, int &c)
{
    const float est = (p - q.rf[j]) * q.inv_vf[j];
    const float fl = floorf(est);
    if (MODE == 0) { if (!(fl >= 0.f) || fl >= (float)q.g[j]) return false; c = (int)fl; return true; }
    const float fr = est - fl;
    const float band = 1e-6f * (fabsf(est) + q.rv_abs[j]) + 1e-6f;
    if (!(fr > band && fr < 1.0f - band)) {
        double cd;
        if (MODE == 1) cd = (double)floorf(__fdiv_rn(__fsub_rn(p, q.rf[j]), q.vf[j]));
        else cd = floor(((double)p - q.r[j]) / q.v[j]);
        if (!(cd >= 0.0) || cd >= (double)q.g[j]) return false;
        c = (int)cd; return true;
    }
    if (!(fl >= 0.f) || fl >= (float)q.g[j]) return false;
    c = (int)fl; return true;
}
template <int MODE>
__global__ void k(const float4* __restrict__ pts, int n, const Prm q, int* out) {
    int i = blockIdx.x*blockDim.x+threadIdx.x; if (i>=n) return;
    float4 v = __ldg(pts+i); int cx,cy,cz; int cell=-1;
    if (axis_cell<MODE>(q,0,v.x,cx) && axis_cell<MODE>(q,1,v.y,cy) && axis_cell<MODE>(q,2,v.z,cz)) cell=(cz*q.g[1]+cy)*q.g[0]+cx;
    out[i]=cell;
}
__global__ void k_all_f64(const float4* __restrict__ pts, int n, const Prm q, int* out) {
    int i = blockIdx.x*blockDim.x+threadIdx.x; if (i>=n) return;
    float4 v = __ldg(pts+i); float p[3]={v.x,v.y,v.z}; int c[3]; bool ok=true;
    for (int j=0;j<3;j++){ double cd=floor(((double)p[j]-q.r[j])/q.v[j]); if(!(cd>=0.0)||cd>=(double)q.g[j]) ok=false; c[j]=(int)cd; }
    out[i]= ok ? (c[2]*q.g[1]+c[1])*q.g[0]+c[0] : -1;
}
int main(){
  const int n=1000000; float4* pts; int* out; cudaMalloc(&pts,n*16); cudaMalloc(&out,n*4);
  float4* h=(float4*)malloc(n*16); srand(2);
  for(int i=0;i<n;i++){ h[i].x=69.12f*rand()/RAND_MAX; h[i].y=-39.68f+79.36f*rand()/RAND_MAX; h[i].z=-3+4.f*rand()/RAND_MAX; h[i].w=0.5f; }
  cudaMemcpy(pts,h,n*16,cudaMemcpyHostToDevice);
  Prm q; double r[3]={0,-39.68,-3}, v[3]={(double)0.16f,(double)0.16f,4.0}; int g[3]={432,496,1};
  for(int j=0;j<3;j++){q.r[j]=r[j];q.v[j]=v[j];q.rf[j]=(float)r[j];q.vf[j]=(float)v[j];q.inv_vf[j]=1.f/q.vf[j];q.rv_abs[j]=(float)(fabs(r[j])/v[j]);q.g[j]=g[j];} q.regime=2;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms; dim3 gr((n+255)/256),b(256);
  #define RUN(name, call) for(int r_=0;r_<3;r_++){call;} cudaEventRecord(e0); for(int r_=0;r_<20;r_++){call;} cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1); printf("%-34s %7.2f us\n", name, ms*1000/20);
  RUN("fp32 estimate only", (k<0><<<gr,b>>>(pts,n,q,out)))
  RUN("estimate + band, f32 exact path", (k<1><<<gr,b>>>(pts,n,q,out)))
  RUN("estimate + band, f64 exact path", (k<2><<<gr,b>>>(pts,n,q,out)))
  RUN("all fp64 divisions", (k_all_f64<<<gr,b>>>(pts,n,q,out)))
  return 0;
}
