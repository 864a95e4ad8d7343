// pipe_probe.cu -- issue-rate microbenchmarks for the instruction mix of the banded fill kernel
// (fp64 add / compare, 64-bit selects, predicated adds).  Question answered: do the fp64 pipe and
// the ALU pipe (FSEL/SEL/LOP3) of an sm_100a SM sub-partition issue independently, i.e. is the fill
// kernel bound by max(fp64, alu) or by their sum?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_probe tools/pipe_probe.cu && tools/pipe_probe
//
// Every kernel runs ITER iterations of an unrolled body of independent operations per thread, with
// WARPS_PER_SM warps resident (default 16 = the fill kernel's occupancy; second argument).  Output:
// warp-instructions per clock per SM for each class in the body.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int N = 8;      // independent chains per thread

// MODE bits: 1 DADD, 2 FSEL pair (64-bit select on a loop-invariant predicate), 4 predicated integer add,
//            8 DSETP (+ predicate consumer folded into the FSEL when both are on), 16 LOP3, 32 IMAD,
//            64 split 64-bit select: low word FSEL (ALU pipe), high word predicated IMAD (replaces bit 2),
//            128 VIADD (add of an immediate; ptxas picks VIADD or IADD3)
template <int MODE>
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed, int iseed, unsigned one)
{
    double a[N], b[N];
    unsigned u[N], w[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed * (i + 3); u[i] = iseed + i + threadIdx.x; w[i] = iseed * 3 + i + 7 * threadIdx.x; }
    const unsigned pv = iseed + threadIdx.x;      // thread-variant, non-zero: keeps the integer work off the uniform datapath
    const double inc = seed * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (MODE & 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[i]) : "d"(inc));
            if (MODE & 64) {
                if (MODE & 8)
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 l0, h0, l1, h1;\n\tsetp.gt.f64 p, %1, %2;\n\t"
                                 "mov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %1;\n\t"
                                 "selp.b32 l0, l0, l1, p;\n\t@!p mad.lo.u32 h0, h1, %3, 0;\n\tmov.b64 %0, {l0, h0};\n\t}"
                                 : "+d"(b[i]) : "d"(a[i]), "d"(inc), "r"(one));
                else
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 l0, h0, l1, h1;\n\tsetp.ne.u32 p, %1, 0;\n\t"
                                 "mov.b64 {l0, h0}, %0;\n\tmov.b64 {l1, h1}, %2;\n\t"
                                 "selp.b32 l0, l0, l1, p;\n\t@!p mad.lo.u32 h0, h1, %3, 0;\n\tmov.b64 %0, {l0, h0};\n\t}"
                                 : "+d"(b[i]) : "r"(pv), "d"(a[i]), "r"(one));
            } else if (MODE & 8) {
                if (MODE & 2)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %0, %1, p;\n\t}"
                                 : "+d"(b[i]) : "d"(a[i]), "d"(inc));
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}"
                                 : "+r"(u[i]) : "d"(a[i]), "d"(inc));
            } else if (MODE & 2) {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tselp.f64 %0, %0, %2, p;\n\t}"
                             : "+d"(b[i]) : "r"(pv), "d"(a[i]));
            }
            if (MODE & 4) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p add.u32 %0, %0, 4;\n\t}" : "+r"(u[i]) : "r"(pv));
            if (MODE & 128) asm volatile("add.u32 %0, %0, 12;" : "+r"(u[i]));
            if (MODE & 16) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(u[i]), "r"(pv));
            if (MODE & 32) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(pv), "r"(u[i]));
        }
    }
    double s = 0;
    unsigned t = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) { s += a[i] + b[i]; t += u[i] + w[i]; }
    if (s == 12345.678 && t == 77) out[0] = s;   // keep the work alive
}

// one warp, one dependent chain: cycles per link
template <int KIND>
__global__ void chain(double* out, int iters, double seed, long long* cycles)
{
    double a = seed + threadIdx.x, b = seed * 3;
    const double inc = seed * 1e-9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (KIND == 0) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a) : "d"(inc));
            if (KIND == 1)      // DADD -> DSETP -> 64-bit select -> (next DADD reads the selected value)
                asm volatile("{\n\t.reg .pred p;\n\tadd.rn.f64 %0, %0, %2;\n\tsetp.gt.f64 p, %0, %1;\n\tselp.f64 %0, %0, %1, p;\n\t}"
                             : "+d"(a) : "d"(b), "d"(inc));
            if (KIND == 2)      // DSETP -> select only
                asm volatile("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %0, %1;\n\tselp.f64 %0, %0, %1, p;\n\t}"
                             : "+d"(a) : "d"(b));
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (a == 12345.678) out[0] = a;
}

void latency(double ghz, double* d_out)
{
    long long* d_c;
    CK(cudaMalloc(&d_c, 8));
    const char* names[3] = {"DADD -> DADD", "DADD -> DSETP -> FSELx2 -> DADD", "DSETP -> FSELx2 -> DSETP"};
    for (int k = 0; k < 3; ++k) {
        const int iters = 2000;
        if (k == 0) chain<0><<<1, 32>>>(d_out, iters, 1.0, d_c);
        if (k == 1) chain<1><<<1, 32>>>(d_out, iters, 1.0, d_c);
        if (k == 2) chain<2><<<1, 32>>>(d_out, iters, 1.0, d_c);
        CK(cudaDeviceSynchronize());
        long long c = 0;
        CK(cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost));
        printf("latency  %-36s %.1f clk per link\n", names[k], (double)c / (iters * 16.0));
    }
}

template <int MODE>
void run(const char* name, int sms, int warps_per_sm, double clock_ghz, double* d_out)
{
    const int iters = 20000;
    const int ctas = sms * warps_per_sm / 8;
    probe<MODE><<<ctas, 256>>>(d_out, 100, 1.0, 1, 1u);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    probe<MODE><<<ctas, 256>>>(d_out, iters, 1.0, 1, 1u);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double per_class = (double)iters * N * warps_per_sm;            // warp-instructions of ONE class per SM
    const double clocks = ms * 1e-3 * clock_ghz * 1e9;
    printf("%-44s %8.3f ms   %.3f warp-instr/clk/SM per class  (%.1f clk per class-instr per SMSP)\n", name, ms,
           per_class / clocks, clocks / (per_class / 4));
}

int main(int argc, char** argv)
{
    const int warps_per_sm = argc > 1 ? atoi(argv[1]) : 16;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int khz = 0;
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (nominal max), %d warps/SM\n", p.name, p.multiProcessorCount, ghz, warps_per_sm);
    double* d_out;
    CK(cudaMalloc(&d_out, 64));
    const int sms = p.multiProcessorCount;
    run<1>("DADD", sms, warps_per_sm, ghz, d_out);
    run<2>("FSELx2 (64-bit select)", sms, warps_per_sm, ghz, d_out);
    run<4>("@p IADD", sms, warps_per_sm, ghz, d_out);
    run<16>("LOP3", sms, warps_per_sm, ghz, d_out);
    run<32>("IMAD", sms, warps_per_sm, ghz, d_out);
    run<1 | 2>("DADD + FSELx2", sms, warps_per_sm, ghz, d_out);
    run<1 | 16>("DADD + LOP3", sms, warps_per_sm, ghz, d_out);
    run<1 | 32>("DADD + IMAD", sms, warps_per_sm, ghz, d_out);
    run<1 | 4>("DADD + @p IADD", sms, warps_per_sm, ghz, d_out);
    run<1 | 8>("DADD + DSETP + @p IADD", sms, warps_per_sm, ghz, d_out);
    run<1 | 8 | 2>("DADD + DSETP + FSELx2", sms, warps_per_sm, ghz, d_out);
    run<1 | 8 | 2 | 4>("DADD + DSETP + FSELx2 + @p IADD", sms, warps_per_sm, ghz, d_out);
    run<2 | 16>("FSELx2 + LOP3", sms, warps_per_sm, ghz, d_out);
    run<2 | 32>("FSELx2 + IMAD", sms, warps_per_sm, ghz, d_out);
    run<128>("add imm (VIADD/IADD3)", sms, warps_per_sm, ghz, d_out);
    run<128 | 32>("add imm + IMAD", sms, warps_per_sm, ghz, d_out);
    run<128 | 16>("add imm + LOP3", sms, warps_per_sm, ghz, d_out);
    run<64>("split select (FSEL + @p IMAD)", sms, warps_per_sm, ghz, d_out);
    run<1 | 64>("DADD + split select", sms, warps_per_sm, ghz, d_out);
    run<1 | 8 | 64>("DADD + DSETP + split select", sms, warps_per_sm, ghz, d_out);
    run<1 | 8 | 64 | 4>("DADD + DSETP + split select + @p IADD", sms, warps_per_sm, ghz, d_out);
    latency(ghz, d_out);
    return 0;
}
