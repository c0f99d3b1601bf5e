#include <cstdio>
#include <cstdint>
struct P { uint32_t sel[16]; };
__device__ __forceinline__ uint32_t packsat(int a, int b, uint32_t c) { uint32_t d; asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) { int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
template<int MODE>
__global__ void __launch_bounds__(256,2) k(const __grid_constant__ P p, int* out, int seed, long long* cyc) {
    int v[8], w[8];
    for (int i=0;i<8;i++){ v[i]=seed+threadIdx.x*7+i; w[i]=seed*3+i; }
    long long t0=clock64();
    #pragma unroll 1
    for (int it=0; it<4096; it++) {
        #pragma unroll
        for (int i=0;i<8;i++) {
            if (MODE==0) v[i] = (int)packsat(v[i], w[i], (uint32_t)v[i]);
            if (MODE==1) v[i] = dp4a_us((uint32_t)v[i], p.sel[i], w[i]);
            if (MODE==2) v[i] = dp4a_us((uint32_t)v[i], 0x01fc01u << (i&1)*8, w[i]);
        }
    }
    long long t1=clock64();
    int acc=0; for(int i=0;i<8;i++) acc^=v[i];
    out[blockIdx.x*256+threadIdx.x]=acc; if(threadIdx.x==0) cyc[blockIdx.x]=t1-t0;
}
int main(){ int* o; long long* c; cudaMalloc(&o,296*256*4); cudaMalloc(&c,296*8); P p; for(int i=0;i<16;i++) p.sel[i]=1u<<(i%4*8);
  static long long h[296];
  #define RUN(M) k<M><<<296,256>>>(p,o,3,c); cudaDeviceSynchronize(); k<M><<<296,256>>>(p,o,3,c); cudaDeviceSynchronize(); cudaMemcpy(h,c,sizeof(h),cudaMemcpyDeviceToHost); { double a=0; for(int i=0;i<296;i++) a+=h[i]; a/=296; printf("mode %d: %.3f warp-inst/clk/SM\n", M, 2.0*8*4096*8/a);} 
  RUN(0) RUN(1) RUN(2) return 0; }
