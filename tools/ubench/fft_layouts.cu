// Micro-benchmark: cuFFT D2Z / Z2D 2048^2 batch B with different half-plane row pitches.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cufft.h>
#define CK(x) do{ auto e_=(x); if(e_){ printf("err %d line %d\n",(int)e_,__LINE__); exit(1);} }while(0)
int main(int argc,char**argv){
  int n = argc>1?atoi(argv[1]):2048; int B = argc>2?atoi(argv[2]):32;
  int pitches[] = {n/2+1, n/2+4, n/2+8, n/2+16, n/2+64};
  for(int pi=0;pi<5;pi++){
    int pc = pitches[pi];
    for(int rp=0; rp<2; rp++){
      int pr = rp? 2*pc : n;   // real pitch: tight, or in-place-compatible padded
      size_t rbytes=(size_t)B*n*pr*8, cbytes=(size_t)B*n*pc*16;
      double *r; cufftDoubleComplex *c; CK(cudaMalloc(&r,rbytes)); CK(cudaMalloc(&c,cbytes));
      CK(cudaMemset(r,0,rbytes)); CK(cudaMemset(c,0,cbytes));
      long long nn[2]={n,n}, ine[2]={n,pr}, one[2]={n,pc};
      cufftHandle f,b; size_t ws;
      CK(cufftCreate(&f)); CK(cufftMakePlanMany64(f,2,nn,ine,1,(long long)n*pr,one,1,(long long)n*pc,CUFFT_D2Z,B,&ws));
      CK(cufftCreate(&b)); CK(cufftMakePlanMany64(b,2,nn,one,1,(long long)n*pc,ine,1,(long long)n*pr,CUFFT_Z2D,B,&ws));
      cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      float msf=0, msb=0;
      for(int it=0; it<3; it++){ CK(cufftExecD2Z(f,r,c)); CK(cufftExecZ2D(b,c,r)); }
      cudaEventRecord(e0); for(int it=0;it<10;it++) CK(cufftExecD2Z(f,r,c)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&msf,e0,e1);
      cudaEventRecord(e0); for(int it=0;it<10;it++) CK(cufftExecZ2D(b,c,r)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&msb,e0,e1);
      double alg = 4.0*8*n*n*B; // 4sN per transform
      printf("n=%d B=%d cpitch=%d rpitch=%d  D2Z %.1f us/map (%.0f GB/s alg)  Z2D %.1f us/map (%.0f GB/s alg) ws=%.0f MB\n", n,B,pc,pr, msf/10/B*1e3, alg/(msf/10*1e-3)/1e9, msb/10/B*1e3, alg/(msb/10*1e-3)/1e9, ws/1e6);
      cufftDestroy(f); cufftDestroy(b); cudaFree(r); cudaFree(c);
    }
  }
  // c2c full plane for reference
  { size_t bytes=(size_t)B*n*n*16; cufftDoubleComplex*a; CK(cudaMalloc(&a,bytes)); CK(cudaMemset(a,0,bytes));
    cufftHandle h; size_t ws; long long nn[2]={n,n}; CK(cufftCreate(&h)); CK(cufftMakePlanMany64(h,2,nn,nullptr,1,0,nullptr,1,0,CUFFT_Z2Z,B,&ws));
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
    for(int it=0;it<3;it++) CK(cufftExecZ2Z(h,a,a,CUFFT_FORWARD));
    cudaEventRecord(e0); for(int it=0;it<10;it++) CK(cufftExecZ2Z(h,a,a,CUFFT_FORWARD)); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms,e0,e1);
    printf("Z2Z in-place n=%d B=%d: %.1f us/map (%.0f GB/s at 4 passes of 16N)\n", n,B, ms/10/B*1e3, 4.0*16*n*n*B/(ms/10*1e-3)/1e9);
  }
  return 0;
}
