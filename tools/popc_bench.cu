// popc_bench.cu -- microbenchmark of the integer pipes the Hamming matcher lives on ("microbenchmark it on the box").
// Prints one JSON object: POPC32 / LOP3 / IADD3 issue rates alone, and 256-bit Hamming distances per second for the two formulations
// of hamming256() (8 XOR + 8 POPC + adds, and a carry-save-adder tree that needs 4 POPC).  Build: tools/build_popc_bench.sh.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 4096, UNROLL = 16;

__global__ void k_popc(uint32_t *out, uint32_t seed)
{
    uint32_t v[UNROLL], acc = 0;
    for (int j = 0; j < UNROLL; j++) v[j] = seed + threadIdx.x * 2654435761u + j * 40503u;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) { v[j] = __popc(v[j]) + v[j]; }      // POPC + IADD (dependent per chain, 16 chains)
    }
    for (int j = 0; j < UNROLL; j++) acc += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_popc_only(uint32_t *out, uint32_t seed)
{
    uint32_t v[UNROLL], acc = 0;
    for (int j = 0; j < UNROLL; j++) v[j] = seed + threadIdx.x * 2654435761u + j * 40503u;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) v[j] = __popc(v[j] | 0x80000000u) | (v[j] << 7);   // POPC + LOP3(shift folded? no): see SASS
    }
    for (int j = 0; j < UNROLL; j++) acc += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
__global__ void k_lop3(uint32_t *out, uint32_t seed)
{
    uint32_t v[UNROLL], acc = 0;
    const uint32_t a = seed * 3u + 1u, b = seed ^ 0x9E3779B9u;
    for (int j = 0; j < UNROLL; j++) v[j] = seed + threadIdx.x * 2654435761u + j * 40503u;
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int j = 0; j < UNROLL; j++) v[j] = (v[j] & a) ^ (b | ~v[(j + 1) % UNROLL]);      // one LOP3 per element
    }
    for (int j = 0; j < UNROLL; j++) acc += v[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__device__ __forceinline__ int ham_plain(const uint32_t (&a)[8], const uint32_t (&b)[8])
{
    int d = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) d += __popc(a[k] ^ b[k]);
    return d;
}
// carry-save adders: sum = x ^ y ^ z, carry = majority -- one LOP3 each
__device__ __forceinline__ void csa(uint32_t x, uint32_t y, uint32_t z, uint32_t &s, uint32_t &c) { s = x ^ y ^ z; c = (x & y) | (z & (x | y)); }
__device__ __forceinline__ int ham_csa(const uint32_t (&a)[8], const uint32_t (&b)[8])
{
    uint32_t x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = a[k] ^ b[k];
    uint32_t s1, c1, s2, c2, s3, c3, s5, c5;
    csa(x[0], x[1], x[2], s1, c1); csa(x[3], x[4], x[5], s2, c2); csa(s1, s2, x[6], s3, c3);
    const uint32_t ones = s3 ^ x[7], c4 = s3 & x[7];
    csa(c1, c2, c3, s5, c5);
    const uint32_t twos = s5 ^ c4, c6 = s5 & c4;
    const uint32_t fours = c5 ^ c6, eights = c5 & c6;
    return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
}
template <int MODE>
__global__ void k_ham(const uint4 *__restrict__ t, int nt, uint32_t *out)
{
    __shared__ uint4 s_t[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_t[i] = t[i % (2 * nt)];
    __syncthreads();
    uint32_t a[8];
    for (int k = 0; k < 8; k++) a[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x;
    int best = 1 << 30;
    for (int rep = 0; rep < 64; rep++) {
#pragma unroll 4
        for (int j = 0; j < 128; j++) {
            const uint4 b0 = s_t[2 * j], b1 = s_t[2 * j + 1];
            const uint32_t b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            const int d = MODE == 0 ? ham_plain(a, b) : ham_csa(a, b);
            best = min(best, (d << 8) | j);
        }
        a[rep & 7] += best;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = best;
}

template <typename F> static float time_ms(F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256;
    uint32_t *out; uint4 *t;
    CK(cudaMalloc(&out, sizeof(uint32_t) * blocks * threads));
    CK(cudaMalloc(&t, sizeof(uint4) * 256));
    CK(cudaMemset(t, 0x5A, sizeof(uint4) * 256));
    const double lanes = (double)blocks * threads;
    const float ms_popc = time_ms([&] { k_popc<<<blocks, threads>>>(out, 1); });
    const float ms_popc2 = time_ms([&] { k_popc_only<<<blocks, threads>>>(out, 1); });
    const float ms_lop3 = time_ms([&] { k_lop3<<<blocks, threads>>>(out, 1); });
    const float ms_h0 = time_ms([&] { k_ham<0><<<blocks, threads>>>(t, 128, out); });
    const float ms_h1 = time_ms([&] { k_ham<1><<<blocks, threads>>>(t, 128, out); });
    CK(cudaGetLastError());
    const double n_ops = lanes * ITERS * UNROLL;
    const double n_ham = lanes * 64.0 * 128.0;
    const double ghz = clk_khz * 1e-6;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_attr\": %.3f,\n", p.name, sms, ghz);
    printf(" \"popc_plus_iadd\": {\"popc_per_s\": %.4e, \"popc_per_clk_per_sm\": %.2f},\n", n_ops / (ms_popc * 1e-3), n_ops / (ms_popc * 1e-3) / (sms * ghz * 1e9));
    printf(" \"popc_plus_lop3\": {\"popc_per_s\": %.4e, \"popc_per_clk_per_sm\": %.2f},\n", n_ops / (ms_popc2 * 1e-3), n_ops / (ms_popc2 * 1e-3) / (sms * ghz * 1e9));
    printf(" \"lop3\": {\"ops_per_s\": %.4e, \"per_clk_per_sm\": %.2f},\n", n_ops / (ms_lop3 * 1e-3), n_ops / (ms_lop3 * 1e-3) / (sms * ghz * 1e9));
    printf(" \"hamming256_plain\": {\"distances_per_s\": %.4e, \"popc32_per_s\": %.4e, \"ms_per_8000x8000\": %.4f},\n", n_ham / (ms_h0 * 1e-3), 8 * n_ham / (ms_h0 * 1e-3),
           64e6 / (n_ham / (ms_h0 * 1e-3)) * 1e3);
    printf(" \"hamming256_csa\": {\"distances_per_s\": %.4e, \"ms_per_8000x8000\": %.4f}}\n", n_ham / (ms_h1 * 1e-3), 64e6 / (n_ham / (ms_h1 * 1e-3)) * 1e3);
    return 0;
}
