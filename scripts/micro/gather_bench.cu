// Microbenchmark: throughput of dependent random 4-byte accesses (L2-resident tables), calibrates the voxelizer model.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); return 1;}}while(0)

__global__ void k_stream(const int* __restrict__ idx, int n, int* out) { int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) out[i]=idx[i]+1; }
__global__ void k_gather1(const int* __restrict__ idx, const int* __restrict__ tab, int n, int* out) { int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) out[i]=tab[idx[i]]; }
__global__ void k_gather2(const int* __restrict__ idx, const int* __restrict__ tab, const int* __restrict__ tab2, int n, int* out) { int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) out[i]=tab2[tab[idx[i]]]; }
__global__ void k_gather3(const int* __restrict__ idx, const int* __restrict__ tab, const int* __restrict__ tab2, int n, int* out) { int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n){int a=tab[idx[i]]; int b=tab2[a]; out[i]=tab[b];} }
__global__ void k_red(const int* __restrict__ idx, int* tab, int n) { int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) atomicAdd(&tab[idx[i]],1); }
__global__ void k_gather_red(const int* __restrict__ idx, const int* __restrict__ tab, int* cnt, int n, int stride) { int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n){int q=tab[idx[i]]; atomicAdd(&cnt[(size_t)q*stride + (i&63)],1);} }
__global__ void k_atom(const int* __restrict__ idx, int* tab, int n, int* out) { int i=blockIdx.x*blockDim.x+threadIdx.x; if(i<n) out[i]=atomicMin(&tab[idx[i]],i); }

int main(){
  const int n=1<<20; const int T=214272; // table entries (cells)
  int *idx,*tab,*tab2,*out,*cnt; 
  CK(cudaMalloc(&idx,n*4)); CK(cudaMalloc(&tab,T*4)); CK(cudaMalloc(&tab2,T*4)); CK(cudaMalloc(&out,n*4)); CK(cudaMalloc(&cnt,(size_t)12000*64*4));
  int *h=(int*)malloc(n*4); 
  // zipf-ish: points concentrated in 11.5k of the cells
  srand(1); int *cells=(int*)malloc(11500*4); for(int i=0;i<11500;i++) cells[i]=rand()%T;
  for(int i=0;i<n;i++) h[i]=cells[rand()%11500];
  CK(cudaMemcpy(idx,h,n*4,cudaMemcpyHostToDevice));
  int *ht=(int*)malloc(T*4); for(int i=0;i<T;i++) ht[i]=rand()%12000; CK(cudaMemcpy(tab,ht,T*4,cudaMemcpyHostToDevice));
  for(int i=0;i<T;i++) ht[i]=rand()%T; CK(cudaMemcpy(tab2,ht,T*4,cudaMemcpyHostToDevice));
  CK(cudaMemset(cnt,0,(size_t)12000*64*4));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  dim3 g((n+255)/256),b(256); float ms;
  #define RUN(name, call) for(int r=0;r<3;r++){call;} cudaEventRecord(e0); for(int r=0;r<20;r++){call;} cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); printf("%-28s %7.2f us\n", name, ms*1000/20);
  RUN("stream copy 1M ints", (k_stream<<<g,b>>>(idx,n,out)))
  RUN("1 random gather", (k_gather1<<<g,b>>>(idx,tab,n,out)))
  RUN("2 dependent gathers", (k_gather2<<<g,b>>>(idx,tab,tab2,n,out)))
  RUN("3 dependent gathers", (k_gather3<<<g,b>>>(idx,tab,tab2,n,out)))
  RUN("RED (no return) random", (k_red<<<g,b>>>(idx,tab2,n)))
  RUN("gather + RED [q][64]", (k_gather_red<<<g,b>>>(idx,tab,cnt,n,64)))
  RUN("ATOM min (return) random", (k_atom<<<g,b>>>(idx,tab2,n,out)))
  return 0;
}
