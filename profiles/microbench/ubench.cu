// profiles/microbench/ubench.cu — instruction-rate probes that shaped the V5 ELA kernel design.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run on one B200.
// Each probe runs a persistent grid (148 SMs x 2 CTAs x 256 threads) of a dependent-chain-free loop and reports
// warp-instructions per clock per SM from clock64() deltas.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

enum Op { OP_IMAD, OP_IADD3, OP_LOP3, OP_SHF, OP_PRMT, OP_DP4A, OP_DP2A, OP_UMULHI, OP_VIMNMX, OP_VIMNMX_RELU,
          OP_MIX_IMAD_IADD, OP_IABS, OP_VABSDIFF4, OP_SHL_ADD, OP_MIX3 };

template <int OP>
__global__ void __launch_bounds__(256, 2) alu_probe(int *out, int seed, long long *cycles)
{
    int v[ILP], w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { v[i] = seed + threadIdx.x * 7 + i; w[i] = seed * 3 + i + 1; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == OP_IMAD) v[i] = v[i] * w[i] + 12345;
            if (OP == OP_IADD3) v[i] = v[i] + w[i] + seed;
            if (OP == OP_LOP3) v[i] = (v[i] ^ w[i]) & seed;
            if (OP == OP_SHF) v[i] = __funnelshift_r(v[i], w[i], 7);
            if (OP == OP_PRMT) v[i] = __byte_perm(v[i], w[i], 0x2143);
            if (OP == OP_DP4A) v[i] = __dp4a(v[i], w[i], v[i]);
            if (OP == OP_DP2A) v[i] = __dp2a_lo(v[i], w[i], v[i]);
            if (OP == OP_UMULHI) v[i] = __umulhi((unsigned)v[i], (unsigned)w[i]) + 1;
            if (OP == OP_VIMNMX) v[i] = max(v[i], w[i]) - 1;
            if (OP == OP_VIMNMX_RELU) v[i] = __vimin_s32_relu(v[i] + 3, w[i]);
            if (OP == OP_MIX_IMAD_IADD) { v[i] = v[i] * w[i] + 77; w[i] = w[i] + v[i] + seed; }
            if (OP == OP_IABS) v[i] = abs(v[i] - w[i]);
            if (OP == OP_VABSDIFF4) v[i] = __vabsdiffu4(v[i], w[i]) + 1;
            if (OP == OP_SHL_ADD) v[i] = (v[i] << 3) + w[i];
            if (OP == OP_MIX3) { v[i] = v[i] * w[i] + 77; w[i] = (w[i] + v[i] + seed) >> 3; }
        }
    }
    long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= v[i] ^ w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// Shared-memory atomic histogram probe: every thread does ITERS atomicAdd(&h[addr], 1) with different address patterns.
// mode 0: all lanes same bin; 1: 16 bins "natural" (hash & 15); 2: spread over 256 bins; 3: 16 bins x 8 lane-copies;
// mode 4: 16 bins x 32 lane-private copies (conflict free); 5: spread 256 x 4 copies
template <int MODE>
__global__ void __launch_bounds__(256, 2) atoms_probe(unsigned *out, unsigned seed, long long *cycles)
{
    __shared__ unsigned h[256 * 32];
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) h[i] = 0;
    __syncthreads();
    unsigned s = seed + threadIdx.x * 2654435761u + blockIdx.x;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        s = s * 1664525u + 1013904223u;
        unsigned r = s >> 16;
        unsigned bin;
        if (MODE == 0) bin = 3;
        if (MODE == 1) bin = r & 15;
        if (MODE == 2) bin = r & 255;
        if (MODE == 3) bin = (r & 15) * 8 + (lane & 7);
        if (MODE == 4) bin = (r & 15) * 32 + lane;
        if (MODE == 5) bin = (r & 255) * 4 + (lane & 3);
        if (MODE == 6) bin = (r & 15) * 4 + (lane & 3);
        if (MODE == 7) bin = (r & 3) * 8 + (lane & 7);      // very peaked content (4 bins) x 8 copies
        if (MODE == 8) bin = 3 * 8 + (lane & 7);            // flat content x 8 copies
        if (MODE == 9) bin = 3 * 32 + lane;                 // flat content x 32 copies
        atomicAdd(&h[bin], 1u);
    }
    long long t1 = clock64();
    __syncthreads();
    unsigned acc = 0;
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) acc += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// match_any aggregated variant: leader lane adds popc(peers)
template <int MODE>
__global__ void __launch_bounds__(256, 2) match_probe(unsigned *out, unsigned seed, long long *cycles)
{
    __shared__ unsigned h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0;
    __syncthreads();
    unsigned s = seed + threadIdx.x * 2654435761u + blockIdx.x;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        s = s * 1664525u + 1013904223u;
        unsigned r = s >> 16;
        unsigned bin = MODE == 0 ? 3 : (MODE == 1 ? (r & 15) : (r & 255));
        unsigned peers = __match_any_sync(0xffffffffu, bin);
        if ((peers & ((1u << lane) - 1)) == 0) atomicAdd(&h[bin], __popc(peers));
    }
    long long t1 = clock64();
    __syncthreads();
    unsigned acc = 0;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) acc += h[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int *d_out;
static long long *d_cyc;
static const int GRID = 148 * 2, BLOCK = 256;

template <typename F>
static int report(const char *name, F launch, double inst_per_thread_iter)
{
    launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    static long long h_cyc[GRID];
    CK(cudaMemcpy(h_cyc, d_cyc, sizeof(h_cyc), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < GRID; i++) avg += (double)h_cyc[i]; avg /= GRID;
    // per SM: 2 CTAs x 8 warps each issue ITERS*inst warp-instructions during ~avg cycles
    double winst_per_clk_sm = 2.0 * 8.0 * ITERS * inst_per_thread_iter / avg;
    printf("%-28s %9.3f ms  avg_cycles %12.0f  warp-inst/clk/SM %6.3f  lane-ops/clk/SM %7.2f  eff_clock_MHz %7.1f\n",
           name, ms, avg, winst_per_clk_sm, winst_per_clk_sm * 32.0, avg / (ms * 1e3));
    return 0;
}

int main()
{
    CK(cudaMalloc(&d_out, GRID * BLOCK * sizeof(int)));
    CK(cudaMalloc(&d_cyc, GRID * sizeof(long long)));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s  SMs %d  clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
#define ALU(OPN, N) report(#OPN, [] { alu_probe<OPN><<<GRID, BLOCK>>>(d_out, 3, d_cyc); }, (double)ILP * (N))
    ALU(OP_IMAD, 1); ALU(OP_IADD3, 1); ALU(OP_LOP3, 1); ALU(OP_SHF, 1); ALU(OP_PRMT, 1); ALU(OP_DP4A, 1); ALU(OP_DP2A, 1);
    ALU(OP_UMULHI, 2); ALU(OP_VIMNMX, 2); ALU(OP_VIMNMX_RELU, 2); ALU(OP_MIX_IMAD_IADD, 2); ALU(OP_IABS, 2);
    ALU(OP_VABSDIFF4, 2); ALU(OP_SHL_ADD, 1); ALU(OP_MIX3, 3);
#define ATM(M) report("atoms_mode" #M, [] { atoms_probe<M><<<GRID, BLOCK>>>((unsigned *)d_out, 5u, d_cyc); }, 1.0)
    ATM(0); ATM(1); ATM(2); ATM(3); ATM(4); ATM(5); ATM(6); ATM(7); ATM(8); ATM(9);
#define MTC(M) report("match_mode" #M, [] { match_probe<M><<<GRID, BLOCK>>>((unsigned *)d_out, 5u, d_cyc); }, 1.0)
    MTC(0); MTC(1); MTC(2);
    printf("note: for atoms/match rows, 'warp-inst/clk/SM' counts one histogram update instruction per iteration\n");
    return 0;
}
