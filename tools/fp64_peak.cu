// FP64 peak microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 .f64),
// plus exp() throughput and rsqrt/div/sqrt dependent-chain latency.  Writes JSON to stdout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){fprintf(stderr,"%s:%d %s\n",__FILE__,__LINE__,cudaGetErrorString(e)); exit(1);} }while(0)

template<int ILP> __global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b){
  double acc[ILP];
  #pragma unroll
  for(int i=0;i<ILP;i++) acc[i]=threadIdx.x*1e-3+i;
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=acc[i];
  if(s==12345.678) out[0]=s;
}

__device__ __forceinline__ void dmma884(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
__device__ __forceinline__ void dmma1688(double* c,const double* a,const double* b){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};"
    :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c,const double* a,const double* b){
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};"
    :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),"d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}
template<int ILP> __global__ void __launch_bounds__(256) k_dmma884(double* out,int iters,double a,double b){
  double c[ILP][2];
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=threadIdx.x;c[i][1]=i;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) dmma884(c[i][0],c[i][1],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1];
  if(s==12345.678) out[0]=s;
}
template<int ILP> __global__ void __launch_bounds__(256) k_dmma1688(double* out,int iters,double a,double b){
  double c[ILP][4]; double af[4]={a,a+1,a+2,a+3}; double bf[2]={b,b+1};
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=threadIdx.x;c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) dmma1688(c[i],af,bf);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  if(s==12345.678) out[0]=s;
}
template<int ILP> __global__ void __launch_bounds__(256) k_dmma16816(double* out,int iters,double a,double b){
  double c[ILP][4]; double af[8]={a,a+1,a+2,a+3,a+4,a+5,a+6,a+7}; double bf[4]={b,b+1,b+2,b+3};
  #pragma unroll
  for(int i=0;i<ILP;i++){c[i][0]=threadIdx.x;c[i][1]=i;c[i][2]=1;c[i][3]=2;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<ILP;i++) dmma16816(c[i],af,bf);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<ILP;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  if(s==12345.678) out[0]=s;
}
__global__ void __launch_bounds__(256) k_exp(double* out,int iters,double a){
  double x0=-1e-3*threadIdx.x, x1=x0-0.5, x2=x0-1.0, x3=x0-1.5, s=0;
  for(int it=0;it<iters;it++){ s+=exp(x0)+exp(x1)+exp(x2)+exp(x3); x0-=a;x1-=a;x2-=a;x3-=a; }
  if(s==12345.678) out[0]=s;
}
// dependent chains, one warp: cycles per op
__global__ void k_lat(long long* cyc,double* out,int iters,double x){
  double v=x; long long t0=clock64();
  for(int i=0;i<iters;i++) v=sqrt(v)+1.5;
  long long t1=clock64(); cyc[0]=(t1-t0)/iters; double s=v;
  v=x; t0=clock64();
  for(int i=0;i<iters;i++) v=1.0/v+1.5;
  t1=clock64(); cyc[1]=(t1-t0)/iters; s+=v;
  v=x; t0=clock64();
  for(int i=0;i<iters;i++) v=rsqrt(v)+1.5;
  t1=clock64(); cyc[2]=(t1-t0)/iters; s+=v;
  v=x; t0=clock64();
  for(int i=0;i<iters;i++) v=fma(v,1.0000001,1.5);
  t1=clock64(); cyc[3]=(t1-t0)/iters; s+=v;
  double c0=x,c1=x; t0=clock64();
  for(int i=0;i<iters;i++) dmma884(c0,c1,1e-3,1e-3);
  t1=clock64(); cyc[4]=(t1-t0)/iters; s+=c0+c1;
  v=x; t0=clock64();
  for(int i=0;i<iters;i++) v=__shfl_sync(0xffffffffu,v,(threadIdx.x+1)&31)+1.0;
  t1=clock64(); cyc[5]=(t1-t0)/iters; s+=v;
  if(s==12345.678) out[0]=s;
}

template<typename F> float timeit(F f,int reps){
  cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms,a,b)); if(ms<best)best=ms; }
  return best;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; double* out; CK(cudaMalloc(&out,64)); long long* cyc; CK(cudaMalloc(&cyc,64));
  int iters=4096; int grid=sms*8; // 8 CTAs x 256 thr = 64 warps/SM
  printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz\":%d",p.name,sms,p.clockRate);
  {float ms=timeit([&]{k_dfma<8><<<grid,256>>>(out,iters,1.0000001,1e-9);},5); printf(",\"dfma_tflops\":%.2f",2.0*grid*256.0*8*iters/ms/1e9);}
  {float ms=timeit([&]{k_dfma<8><<<sms*4,256>>>(out,iters,1.0000001,1e-9);},5); printf(",\"dfma_tflops_occ32w\":%.2f",2.0*sms*4*256.0*8*iters/ms/1e9);}
  {float ms=timeit([&]{k_dmma884<8><<<grid,256>>>(out,iters,1e-3,1e-3);},5); printf(",\"dmma884_tflops\":%.2f",2.0*256*(grid*8.0)*8*iters/ms/1e9);}
  {float ms=timeit([&]{k_dmma884<8><<<sms,128>>>(out,iters,1e-3,1e-3);},5); printf(",\"dmma884_tflops_4warps\":%.2f",2.0*256*(sms*4.0)*8*iters/ms/1e9);}
  {float ms=timeit([&]{k_dmma884<8><<<sms*2,256>>>(out,iters,1e-3,1e-3);},5); printf(",\"dmma884_tflops_16warps\":%.2f",2.0*256*(sms*16.0)*8*iters/ms/1e9);}
  {float ms=timeit([&]{k_dmma1688<4><<<grid,256>>>(out,iters,1e-3,1e-3);},5); printf(",\"dmma1688_tflops\":%.2f",2.0*1024*(grid*8.0)*4*iters/ms/1e9);}
  {float ms=timeit([&]{k_dmma16816<4><<<grid,256>>>(out,iters,1e-3,1e-3);},5); printf(",\"dmma16816_tflops\":%.2f",2.0*2048*(grid*8.0)*4*iters/ms/1e9);}
  {float ms=timeit([&]{k_exp<<<grid,256>>>(out,iters,1e-4);},5); printf(",\"exp_gops\":%.1f",grid*256.0*4*iters/ms/1e6);}
  // sustained: the same kernels back to back for ~3 s each (power / clock settle), one event pair around the whole run
  {
    auto sustained=[&](auto launch,double flop_per_launch){
      cudaEvent_t a,b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      launch(); CK(cudaDeviceSynchronize());
      float one=timeit(launch,2); int reps=(int)(3000.0f/one)+1;
      CK(cudaEventRecord(a)); for(int r=0;r<reps;r++) launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms,a,b)); return flop_per_launch*reps/ms/1e9;
    };
    printf(",\"dfma_tflops_sustained\":%.2f",sustained([&]{k_dfma<8><<<grid,256>>>(out,iters,1.0000001,1e-9);},2.0*grid*256.0*8*iters));
    printf(",\"dmma884_tflops_sustained\":%.2f",sustained([&]{k_dmma884<8><<<grid,256>>>(out,iters,1e-3,1e-3);},2.0*256*(grid*8.0)*8*iters));
  }
  k_lat<<<1,32>>>(cyc,out,2000,1.7); CK(cudaDeviceSynchronize()); long long h[6]; CK(cudaMemcpy(h,cyc,48,cudaMemcpyDeviceToHost));
  printf(",\"lat_cycles\":{\"sqrt_add\":%lld,\"div_add\":%lld,\"rsqrt_add\":%lld,\"dfma\":%lld,\"dmma884\":%lld,\"shfl64_add\":%lld}}\n",h[0],h[1],h[2],h[3],h[4],h[5]);
  return 0;
}
