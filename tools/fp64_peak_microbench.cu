// FP64 peak microbenchmarks: DFMA pipe vs DMMA (mma.sync f64) on sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

template<int NACC>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b){
  double acc[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++) acc[i]=threadIdx.x*1e-3+i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) acc[i]=fma(acc[i],a,b);
  }
  double s=0; 
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=acc[i];
  if(s==123.456) out[0]=s;
}

__device__ __forceinline__ void dmma884(double& c0,double& c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}
template<int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b){
  double c0[NACC], c1[NACC];
  #pragma unroll
  for(int i=0;i<NACC;i++){ c0[i]=threadIdx.x*1e-3+i; c1[i]=i; }
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma884(c0[i],c1[i],a,b);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c0[i]+c1[i];
  if(s==123.456) out[0]=s;
}
// m16n8k8: A 4 regs, B 2 regs, C 4 regs
__device__ __forceinline__ void dmma1688(double* c,const double* a,const double* b){
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
   :"+d"(c[0]),"+d"(c[1]),"+d"(c[2]),"+d"(c[3]):"d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(b[0]),"d"(b[1]));
}
template<int NACC>
__global__ void __launch_bounds__(256) k_dmma1688(double* out, int iters, double a, double b){
  double c[NACC][4];
  #pragma unroll
  for(int i=0;i<NACC;i++){ c[i][0]=threadIdx.x*1e-3+i; c[i][1]=i; c[i][2]=1; c[i][3]=2;}
  double av[4]={a,a+1,a+2,a+3}, bv[2]={b,b+1};
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NACC;i++) dmma1688(c[i],av,bv);
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
  if(s==123.456) out[0]=s;
}
// mixed: NA dmma accumulators + NF dfma accumulators interleaved
template<int NA,int NF>
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b){
  double c0[NA], c1[NA], f[NF];
  #pragma unroll
  for(int i=0;i<NA;i++){ c0[i]=threadIdx.x*1e-3+i; c1[i]=i; }
  #pragma unroll
  for(int i=0;i<NF;i++) f[i]=threadIdx.x+i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int i=0;i<NA;i++){ dmma884(c0[i],c1[i],a,b);
      #pragma unroll
      for(int j=0;j<NF/NA;j++) f[i*(NF/NA)+j]=fma(f[i*(NF/NA)+j],a,b);
    }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<NA;i++) s+=c0[i]+c1[i];
  #pragma unroll
  for(int i=0;i<NF;i++) s+=f[i];
  if(s==123.456) out[0]=s;
}

template<typename F> float timeit(F f, int reps){
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for(int r=0;r<reps;r++) f();
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); return ms/reps;
}
int main(int argc,char**argv){
  int iters = argc>1? atoi(argv[1]) : 20000;
  int reps  = argc>2? atoi(argv[2]) : 5;
  double* out; CK(cudaMalloc(&out,8));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  int sms=p.multiProcessorCount; printf("SMs %d clock %d kHz\n", sms, p.clockRate);
  for(int bps : {1,2,4,8}) {
    int grid=sms*bps;
    {float ms=timeit([&]{k_dfma<16><<<grid,256>>>(out,iters,1.0000001,1e-9);},reps);
     double fl=2.0*16*iters*256.0*grid; printf("dfma16   bps=%d  %.3f ms  %.2f TF\n",bps,ms,fl/ms*1e-9);}
    {float ms=timeit([&]{k_dmma<8><<<grid,256>>>(out,iters,1.0000001,1e-9);},reps);
     double fl=2.0*8*8*4*8*iters*8.0*grid; printf("dmma884x8 bps=%d  %.3f ms  %.2f TF\n",bps,ms,fl/ms*1e-9);}
    {float ms=timeit([&]{k_dmma<16><<<grid,256>>>(out,iters,1.0000001,1e-9);},reps);
     double fl=2.0*8*8*4*16*iters*8.0*grid; printf("dmma884x16 bps=%d  %.3f ms  %.2f TF\n",bps,ms,fl/ms*1e-9);}
    {float ms=timeit([&]{k_dmma1688<8><<<grid,256>>>(out,iters,1.0000001,1e-9);},reps);
     double fl=2.0*16*8*8*8*iters*8.0*grid; printf("dmma1688x8 bps=%d  %.3f ms  %.2f TF\n",bps,ms,fl/ms*1e-9);}
    {float ms=timeit([&]{k_mixed<8,8><<<grid,256>>>(out,iters,1.0000001,1e-9);},reps);
     double fl=(2.0*8*8*4*8*8.0 + 2.0*8*256.0)*iters*grid; printf("mixed 8dmma+8dfma bps=%d  %.3f ms  %.2f TF\n",bps,ms,fl/ms*1e-9);}
    {float ms=timeit([&]{k_mixed<4,16><<<grid,256>>>(out,iters,1.0000001,1e-9);},reps);
     double fl=(2.0*8*8*4*4*8.0 + 2.0*16*256.0)*iters*grid; printf("mixed 4dmma+16dfma bps=%d  %.3f ms  %.2f TF\n",bps,ms,fl/ms*1e-9);}
  }
  // sustained run: ~3 s of dfma and dmma
  {int grid=sms*4; float ms=timeit([&]{k_dfma<16><<<grid,256>>>(out,iters*10,1.0000001,1e-9);},20);
   double fl=2.0*16*iters*10*256.0*grid; printf("SUSTAINED dfma16 %.3f ms/launch  %.2f TF\n",ms,fl/ms*1e-9);}
  {int grid=sms*4; float ms=timeit([&]{k_dmma<8><<<grid,256>>>(out,iters*10,1.0000001,1e-9);},20);
   double fl=2.0*8*8*4*8*iters*10*8.0*grid; printf("SUSTAINED dmma884x8 %.3f ms/launch  %.2f TF\n",ms,fl/ms*1e-9);}
  return 0;
}
