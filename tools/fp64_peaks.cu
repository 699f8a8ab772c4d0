// Micro-benchmark: FP64 DFMA / DMMA issue peaks and HBM write/read bandwidth on B200.
// Numbers feed DESIGN.md (secondary FP64 bound of the dense-Hessian kernel).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__global__ void dfma_kernel(double* out, int iters, double a, double b){
    double acc[16];
#pragma unroll
    for(int i=0;i<16;i++) acc[i]=threadIdx.x*1e-3+i;
    for(int it=0; it<iters; it++){
#pragma unroll
        for(int i=0;i<16;i++) acc[i]=fma(acc[i],a,b);
    }
    double s=0;
#pragma unroll
    for(int i=0;i<16;i++) s+=acc[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

__device__ __forceinline__ void dmma884(double &c0,double &c1,double a,double b){
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmma884_kernel(double* out, int iters, double a, double b){
    double c[16];
#pragma unroll
    for(int i=0;i<16;i++) c[i]=threadIdx.x*1e-3+i;
    for(int it=0; it<iters; it++){
#pragma unroll
        for(int i=0;i<8;i++) dmma884(c[2*i],c[2*i+1],a,b);
    }
    double s=0;
#pragma unroll
    for(int i=0;i<16;i++) s+=c[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
// m16n8k16 f64 (sm_90+): A 16x16 (8 regs), B 16x8 (4 regs), C 16x8 (4 regs)
__device__ __forceinline__ void dmma16816(double* c,const double* a,const double* b){
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]),"d"(a[1]),"d"(a[2]),"d"(a[3]),"d"(a[4]),"d"(a[5]),"d"(a[6]),"d"(a[7]),
                   "d"(b[0]),"d"(b[1]),"d"(b[2]),"d"(b[3]));
}
__global__ void dmma16816_kernel(double* out, int iters, double av, double bv){
    double c[4][4]; double a[8]; double b[4];
#pragma unroll
    for(int i=0;i<8;i++) a[i]=av+i*1e-9;
#pragma unroll
    for(int i=0;i<4;i++) b[i]=bv+i*1e-9;
#pragma unroll
    for(int j=0;j<4;j++)
#pragma unroll
    for(int i=0;i<4;i++) c[j][i]=threadIdx.x*1e-3+i+j;
    for(int it=0; it<iters; it++){
#pragma unroll
        for(int j=0;j<4;j++) dmma16816(c[j],a,b);
    }
    double s=0;
#pragma unroll
    for(int j=0;j<4;j++)
#pragma unroll
    for(int i=0;i<4;i++) s+=c[j][i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

__global__ void write_default(double2* p, size_t n2){
    size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
    double2 v=make_double2(1.0,2.0);
    for(;i<n2;i+=stride) p[i]=v;
}
__global__ void write_cs(double2* p, size_t n2){
    size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
    double2 v=make_double2(1.0,2.0);
    for(;i<n2;i+=stride) __stcs(p+i,v);
}
// each lane writes 32B-sector chunks in a strided pattern: lane l of warp writes 16B at row (l/4), 64B contiguous per row
__global__ void write_rows64(double* p, size_t ld, size_t rows){
    // grid-stride over 8-row x 8-col (64B) tiles; tile t -> row block (t / (ld/8)), col block (t % (ld/8))
    size_t warp=(blockIdx.x*(size_t)blockDim.x+threadIdx.x)>>5; size_t nw=((size_t)gridDim.x*blockDim.x)>>5;
    int lane=threadIdx.x&31; size_t cb=ld/8; size_t ntiles=(rows/8)*cb;
    double2 v=make_double2(1.0,2.0);
    for(size_t t=warp;t<ntiles;t+=nw){
        size_t r=(t/cb)*8+(lane>>2), c=(t%cb)*8+(lane&3)*2;
        __stcs((double2*)(p+r*ld+c),v);
    }
}
__global__ void read_sum(const double2* p, size_t n2, double* out){
    size_t i=blockIdx.x*(size_t)blockDim.x+threadIdx.x; size_t stride=(size_t)gridDim.x*blockDim.x;
    double s=0;
    for(;i<n2;i+=stride){ double2 v=__ldcs(p+i); s+=v.x+v.y; }
    if(s==123.456) out[0]=s;
}

template<class F> float timeit(F f,int reps){
    cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f(); CK(cudaDeviceSynchronize());
    float best=1e30f;
    for(int r=0;r<reps;r++){ CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms; }
    return best;
}
int main(){
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr,0));
    printf("device %s SMs %d clock %d kHz\n",pr.name,pr.multiProcessorCount,pr.clockRate);
    int nsm=pr.multiProcessorCount;
    double* out; CK(cudaMalloc(&out,sizeof(double)*nsm*8*1024));
    int iters=4096;
    for(int bpsm=1;bpsm<=4;bpsm*=2){
        int grid=nsm*bpsm, blk=512;
        float ms=timeit([&]{dfma_kernel<<<grid,blk>>>(out,iters,1.0000001,1e-9);},5);
        double fl=2.0*16*iters*(double)grid*blk;
        printf("DFMA      grid=%d blk=%d: %.3f ms  %.2f TFLOP/s\n",grid,blk,ms,fl/ms*1e-9);
        ms=timeit([&]{dmma884_kernel<<<grid,blk>>>(out,iters,1.0000001,1e-9);},5);
        fl=2.0*8*256*iters*(double)grid*blk/32;
        printf("DMMA884   grid=%d blk=%d: %.3f ms  %.2f TFLOP/s\n",grid,blk,ms,fl/ms*1e-9);
        ms=timeit([&]{dmma16816_kernel<<<grid,blk>>>(out,iters,1.0000001,1e-9);},5);
        fl=2.0*4*(16*8*16)*iters*(double)grid*blk/32;
        printf("DMMA16816 grid=%d blk=%d: %.3f ms  %.2f TFLOP/s\n",grid,blk,ms,fl/ms*1e-9);
    }
    size_t bytes=(size_t)8<<30; double* buf; CK(cudaMalloc(&buf,bytes)); size_t n2=bytes/16;
    for(int bpsm=2;bpsm<=16;bpsm*=2){
        int grid=nsm*bpsm, blk=512;
        float ms=timeit([&]{write_default<<<grid,blk>>>((double2*)buf,n2);},5);
        printf("write default grid=%d: %.3f ms %.1f GB/s\n",grid,ms,bytes/ms*1e-6);
        ms=timeit([&]{write_cs<<<grid,blk>>>((double2*)buf,n2);},5);
        printf("write .cs     grid=%d: %.3f ms %.1f GB/s\n",grid,ms,bytes/ms*1e-6);
        size_t ld=32768, rows=bytes/8/ld;
        ms=timeit([&]{write_rows64<<<grid,blk>>>(buf,ld,rows);},5);
        printf("write rows64  grid=%d: %.3f ms %.1f GB/s\n",grid,ms,bytes/ms*1e-6);
        ms=timeit([&]{read_sum<<<grid,blk>>>((const double2*)buf,n2,out);},5);
        printf("read  .cs     grid=%d: %.3f ms %.1f GB/s\n",grid,ms,bytes/ms*1e-6);
    }
    float ms=timeit([&]{CK(cudaMemsetAsync(buf,0,bytes));},5);
    printf("cudaMemset: %.3f ms %.1f GB/s\n",ms,bytes/ms*1e-6);
    double* buf2; CK(cudaMalloc(&buf2,bytes/2));
    ms=timeit([&]{CK(cudaMemcpyAsync(buf2,buf,bytes/2,cudaMemcpyDeviceToDevice));},5);
    printf("cudaMemcpy D2D (r+w bytes): %.3f ms %.1f GB/s\n",ms,bytes/ms*1e-6);
    // pinned host D2H bandwidth
    void* h; size_t hb=(size_t)2<<30; CK(cudaMallocHost(&h,hb));
    ms=timeit([&]{CK(cudaMemcpyAsync(h,buf,hb,cudaMemcpyDeviceToHost));},3);
    printf("D2H pinned: %.3f ms %.1f GB/s\n",ms,hb/ms*1e-6);
    ms=timeit([&]{CK(cudaMemcpyAsync(buf,h,hb,cudaMemcpyHostToDevice));},3);
    printf("H2D pinned: %.3f ms %.1f GB/s\n",ms,hb/ms*1e-6);
    return 0;
}
