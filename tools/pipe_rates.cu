// Developer tool: reciprocal throughput of the instruction kinds the CLAHE cell loop is built from, on the GPU it runs on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_rates tools/pipe_rates.cu && /tmp/pipe_rates
// Every test runs 8 independent dependency chains per thread, 8 warps per SM sub-partition (1024 threads, one CTA per SM),
// and prints cycles per warp-instruction per sub-partition (1.0 = the issue limit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

__device__ __forceinline__ uint64_t pk(float a, float b) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) rate_kernel(unsigned long long* out, float seed, uint32_t useed) {
    extern __shared__ uint32_t smem[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 16384; i += 1024) smem[i] = 0;
    __syncthreads();
    float f[8];
    uint32_t u[8];
    uint64_t d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        f[i] = seed + i + tid * 1e-3f;
        u[i] = useed * (i + 1) + tid;
        d[i] = pk(f[i], f[i] + 1.0f);
    }
    if (OP == 12) {   // subnormal inputs
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = ((uint64_t)(u[i] & 255u) << 32) | (u[i] & 255u);
    }
    const float c = seed * 0.999f;
    const uint64_t c2 = pk(c, c);
    const uint32_t lane4 = (tid & 31) * 4, lane8 = (tid & 15) * 8;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (OP == 0) {
#define X(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 1) {
#define X(i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 2 || OP == 12) {
#define X(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c2));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 3) {
#define X(i) asm volatile("add.rn.ftz.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c2));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 4) {
#define X(i) asm volatile("mul.lo.u32 %0, %0, 65537;" : "+r"(u[i]));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 5) {
#define X(i) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(d[i]) : "r"((uint32_t)d[i]), "r"(useed));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 6) {
#define X(i) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(u[i]) : "r"(useed));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 7) {
#define X(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(useed), "r"(lane4));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 8) {   // alternate FMUL2 (fma pipe) and PRMT (alu pipe)
#define X(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c2)); asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(u[i]) : "r"(useed));
            REP8(X)
#undef X
        } else if (OP == 9) {   // conflict-free 8-byte gathers (16 replicas of 8 bytes per 256-byte row)
#define X(i) { uint32_t a = ((u[i] & 63u) << 8) + lane8, lo, hi; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a)); u[i] += lo + hi + 1; }
            REP8(X) REP8(X)
#undef X
        } else if (OP == 10) {  // conflict-free shared increments (lane columns)
#define X(i) { uint32_t a = ((u[i] & 63u) << 8) + lane4; asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory"); u[i] += 7; }
            REP8(X) REP8(X)
#undef X
        } else if (OP == 11) {  // scalar FMUL alternating with scalar FADD
#define X(i) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[(i + 4) & 7]) : "f"(c));
            REP8(X)
#undef X
        } else if (OP == 13) {  // FMUL2 + scalar FADD + PRMT + LOP3 (a blend-like mix)
#define X(i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(c2)); asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(u[i]) : "r"(useed)); \
             asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[(i + 4) & 7]) : "r"(useed), "r"(lane4));
            REP8(X)
#undef X
        } else if (OP == 14) {  // shift left by 16 (how ptxas emits it is the question: SHF / IMAD.SHL / IMAD.U32)
#define X(i) asm volatile("shl.b32 %0, %0, 1;" : "+r"(u[i]));
            REP8(X) REP8(X)
#undef X
        } else if (OP == 15) {  // 4-byte gathers, 32 replicas of 4 bytes per 128-byte row: one wavefront per warp
#define X(i) { uint32_t a = ((u[i] & 127u) << 7) + lane4, lo; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(a)); u[i] += lo + 1; }
            REP8(X) REP8(X)
#undef X
        } else if (OP == 16) {  // gathers and increments interleaved
#define X(i) { uint32_t a = ((u[i] & 31u) << 8) + lane8, lo, hi; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a)); u[i] += lo + hi + 1; \
               uint32_t b = 8192 + ((u[i] & 31u) << 8) + lane4; asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(b) : "memory"); }
            REP8(X)
#undef X
        } else if (OP == 17) {  // I2F of a byte
#define X(i) { float r; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(r) : "r"(u[i] & 255u)); u[i] += __float_as_uint(r); }
            REP8(X) REP8(X)
#undef X
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    uint32_t uacc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc += f[i];
        uacc += u[i] + (uint32_t)d[i] + (uint32_t)(d[i] >> 32);
    }
    if (tid == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    if (acc == 123.456f && uacc == 77) out[1] = 1;
}

template <int OP>
void run(const char* name, int per_iter, unsigned long long* d_out) {
    cudaFuncSetAttribute(rate_kernel<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int k = 0; k < 2; ++k) rate_kernel<OP><<<148, 1024, 65536>>>(d_out, 1.0001f, 0x9e3779b9u);
    cudaDeviceSynchronize();
    unsigned long long cyc = 0;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    const double per = (double)cyc / ((double)ITERS * per_iter * 8.0);   // 8 warps per sub-partition
    printf("%-44s %6.2f cycles per warp-instruction per sub-partition (%s)\n", name, per, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    unsigned long long* d_out;
    cudaMalloc(&d_out, 64);
    cudaMemset(d_out, 0, 64);
    run<0>("FMUL scalar", 16, d_out);
    run<1>("FADD scalar", 16, d_out);
    run<11>("FMUL + FADD scalar alternating", 16, d_out);
    run<2>("FMUL2 (mul.rn.f32x2)", 16, d_out);
    run<12>("FMUL2 with subnormal inputs", 16, d_out);
    run<3>("FADD2.FTZ", 16, d_out);
    run<4>("IMAD (mul.lo.u32 by 65537)", 16, d_out);
    run<5>("IMAD.WIDE.U32 (register multiplier)", 16, d_out);
    run<14>("shl 1", 16, d_out);
    run<6>("PRMT", 16, d_out);
    run<7>("LOP3", 16, d_out);
    run<8>("FMUL2 + PRMT alternating", 16, d_out);
    run<13>("FMUL2 + PRMT + FADD + LOP3", 32, d_out);
    run<17>("I2F.U32 + LOP + IADD (3 instr)", 16, d_out);
    run<9>("LDS.64 gather, 16 replicas (+2 int ops)", 16, d_out);
    run<15>("LDS.32 gather, 32 replicas (+2 int ops)", 16, d_out);
    run<10>("ATOMS inc, lane columns (+2 int ops)", 16, d_out);
    run<16>("LDS.64 + ATOMS pairs (per pair)", 8, d_out);
    return 0;
}
