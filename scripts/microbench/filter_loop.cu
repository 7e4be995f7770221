// Microbenchmark (design experiment, not product): the conservative sphere-filter loop of rt_trace.cuh in isolation.
//   scalar<NP>: the FFMA form (8 spheres per group, NP rays per lane, broadcast LDS.128 of {c, w})
//   packed<NP>: Blackwell packed FP32 (fma.rn.f32x2 -> SASS FFMA2): two SPHERES per instruction, the list stored as
//               sphere pairs {x0,x1,y0,y1}{z0,z1,-w0,-w1}; ray constants duplicated into register pairs
//   chains    : FFMA vs FFMA2 dependent-chain throughput (is FFMA2 full rate?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o filter_loop filter_loop.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct F4 { float x, y, z, w; };
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int NP, int CTAS>
__global__ void __launch_bounds__(256, CTAS) scalar(float* out, const F4* gsph, int n, int iters)
{
    extern __shared__ F4 s_hot[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_hot[i] = gsph[i];
    __syncthreads();
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float ox[NP], oy[NP], oz[NP], dx[NP], dy[NP], dz[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { ox[p] = 0.01f * ((tid + p) & 63); oy[p] = 0.02f * ((tid >> 6) & 63); oz[p] = 0.5f + p; dx[p] = 0.6f; dy[p] = 0.0f; dz[p] = -0.8f; }
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float od[NP], kray[NP], px[NP], py[NP], pz[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            od[p] = ox[p] * dx[p] + oy[p] * dy[p] + oz[p] * dz[p]; kray[p] = (ox[p] * ox[p] + oy[p] * oy[p] + oz[p] * oz[p]);
            px[p] = -2.f * ox[p]; py[p] = -2.f * oy[p]; pz[p] = -2.f * oz[p];
        }
        for (int i = 0; i < n; i += 8) {
            float v[NP][8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                F4 s = s_hot[i + k];
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    float hb = fmaf(-s.x, dx[p], fmaf(-s.y, dy[p], fmaf(-s.z, dz[p], od[p])));
                    float t  = fmaf(s.x, px[p], fmaf(s.y, py[p], fmaf(s.z, pz[p], s.w)));
                    v[p][k] = fmaf(hb, hb, -t) - kray[p];
                }
            }
            float m = v[0][0];
#pragma unroll
            for (int p = 0; p < NP; ++p)
#pragma unroll
                for (int k = 0; k < 8; ++k) m = fmaxf(m, v[p][k]);
            if (m >= 0.f) {
#pragma unroll
                for (int p = 0; p < NP; ++p)
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (v[p][k] >= 0.f) acc += sqrtf(v[p][k]) + (float)(i + k);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) { ox[p] += 1e-6f + acc * 1e-12f; dx[p] -= acc * 1e-12f; }
    }
    out[tid] = acc;
}

// list layout: pair j -> s_hot[2j] = {x0, x1, y0, y1}, s_hot[2j+1] = {z0, z1, -w0, -w1}
template <int NP, int CTAS>
__global__ void __launch_bounds__(256, CTAS) packed(float* out, const F4* gpairs, int n, int iters)
{
    extern __shared__ F4 s_hot[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_hot[i] = gpairs[i];
    __syncthreads();
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    float ox[NP], oy[NP], oz[NP], dx[NP], dy[NP], dz[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) { ox[p] = 0.01f * ((tid + p) & 63); oy[p] = 0.02f * ((tid >> 6) & 63); oz[p] = 0.5f + p; dx[p] = 0.6f; dy[p] = 0.0f; dz[p] = -0.8f; }
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        u64   NDX[NP], NDY[NP], NDZ[NP], OD[NP], PX[NP], PY[NP], PZ[NP];
        float kray[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float od = ox[p] * dx[p] + oy[p] * dy[p] + oz[p] * dz[p];
            kray[p] = (ox[p] * ox[p] + oy[p] * oy[p] + oz[p] * oz[p]);
            NDX[p] = pk(-dx[p], -dx[p]); NDY[p] = pk(-dy[p], -dy[p]); NDZ[p] = pk(-dz[p], -dz[p]); OD[p] = pk(od, od);
            PX[p] = pk(2.f * ox[p], 2.f * ox[p]); PY[p] = pk(2.f * oy[p], 2.f * oy[p]); PZ[p] = pk(2.f * oz[p], 2.f * oz[p]);
        }
        for (int i = 0; i < n; i += 8) {            // 8 spheres = 4 pairs = 8 float4
            float u[NP][8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const ulonglong2 A = *reinterpret_cast<const ulonglong2*>(&s_hot[i + 2 * j]);        // X pair, Y pair
                const ulonglong2 B = *reinterpret_cast<const ulonglong2*>(&s_hot[i + 2 * j + 1]);    // Z pair, -W pair
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    u64 hb = fma2(A.x, NDX[p], fma2(A.y, NDY[p], fma2(B.x, NDZ[p], OD[p])));
                    u64 nt = fma2(A.x, PX[p], fma2(A.y, PY[p], fma2(B.x, PZ[p], B.y)));
                    u64 uu = fma2(hb, hb, nt);
                    upk(uu, u[p][2 * j], u[p][2 * j + 1]);
                }
            }
            bool any = false;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                float m = u[p][0];
#pragma unroll
                for (int k = 1; k < 8; ++k) m = fmaxf(m, u[p][k]);
                any = any || (m >= kray[p]);
            }
            if (any) {
#pragma unroll
                for (int p = 0; p < NP; ++p)
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (u[p][k] >= kray[p]) acc += sqrtf(u[p][k]) + (float)(i + k);
            }
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) { ox[p] += 1e-6f + acc * 1e-12f; dx[p] -= acc * 1e-12f; }
    }
    out[tid] = acc;
}

template <bool PACKED>
__global__ void __launch_bounds__(256, 8) chains(float* out, int iters)
{
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float b = 1.0000001f, c = 1e-7f;
    if (PACKED) {
        u64 A[8]; const u64 B = pk(b, b), C = pk(c, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) A[i] = pk(a[2 * i], a[2 * i + 1]);
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) A[i] = fma2(A[i], B, C);
#pragma unroll
        for (int i = 0; i < 8; ++i) upk(A[i], a[2 * i], a[2 * i + 1]);
    } else {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
static void run(const char* name, K kern, int grid, int block, size_t smem, float* out, const F4* d, int n, int iters, int np, int sms)
{
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        kern<<<grid, block, smem>>>(out, d, n, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) best = ms < best ? ms : best;
    }
    double tests = (double)grid * block * iters * n * np;            // lane ray-sphere tests
    printf("%-14s grid %4d  %.3f ms  %7.1f G lane-tests/s  %.3f clk/warp-ray-sphere/SM  (%.0f%% of the 17-flop FFMA peak)  %s\n", name, grid, best,
           tests / best / 1e6, 1.965e9 * (best * 1e-3) / (tests / 32 / sms), 100.0 * 2.125 / (1.965e9 * (best * 1e-3) / (tests / 32 / sms)),
           cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 1000, iters = argc > 2 ? atoi(argv[2]) : 100;
    n = (n + 7) / 8 * 8;
    std::vector<F4> h(n), hp(n);
    for (int i = 0; i < n; ++i) h[i] = {10.f + (i % 37) * 0.5f, -0.3f, -5.f - (i / 37) * 0.5f, 1e3f};     // all miss
    for (int j = 0; j < n / 2; ++j) {
        hp[2 * j]     = {h[2 * j].x, h[2 * j + 1].x, h[2 * j].y, h[2 * j + 1].y};
        hp[2 * j + 1] = {h[2 * j].z, h[2 * j + 1].z, -h[2 * j].w, -h[2 * j + 1].w};
    }
    F4 *d, *dp; cudaMalloc(&d, n * sizeof(F4)); cudaMemcpy(d, h.data(), n * sizeof(F4), cudaMemcpyHostToDevice);
    cudaMalloc(&dp, n * sizeof(F4)); cudaMemcpy(dp, hp.data(), n * sizeof(F4), cudaMemcpyHostToDevice);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * sizeof(float));
    size_t smem = n * sizeof(F4);
    run("scalar NP1 x4", scalar<1, 4>, sms * 4, 256, smem, out, d, n, iters, 1, sms);
    run("scalar NP1 x3", scalar<1, 3>, sms * 3, 256, smem, out, d, n, iters, 1, sms);
    run("scalar NP2 x2", scalar<2, 2>, sms * 2, 256, smem, out, d, n, iters, 2, sms);
    run("scalar NP2 x3", scalar<2, 3>, sms * 3, 256, smem, out, d, n, iters, 2, sms);
    run("scalar NP4 x1", scalar<4, 1>, sms * 1, 256, smem, out, d, n, iters, 4, sms);
    run("scalar NP4 x2", scalar<4, 2>, sms * 2, 256, smem, out, d, n, iters, 4, sms);
    run("packed NP1 x4", packed<1, 4>, sms * 4, 256, smem, out, dp, n, iters, 1, sms);
    run("packed NP1 x3", packed<1, 3>, sms * 3, 256, smem, out, dp, n, iters, 1, sms);
    run("packed NP2 x2", packed<2, 2>, sms * 2, 256, smem, out, dp, n, iters, 2, sms);
    run("packed NP2 x3", packed<2, 3>, sms * 3, 256, smem, out, dp, n, iters, 2, sms);
    run("packed NP4 x1", packed<4, 1>, sms * 1, 256, smem, out, dp, n, iters, 4, sms);
    run("packed NP4 x2", packed<4, 2>, sms * 2, 256, smem, out, dp, n, iters, 4, sms);
    // chains
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int packedm = 0; packedm < 2; ++packedm) {
        float best = 1e30f; const int it = 16384, grid = sms * 8;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (packedm) chains<true><<<grid, 256>>>(out, it); else chains<false><<<grid, 256>>>(out, it);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        printf("%s chains: %.3f ms  %.1f TFLOP/s\n", packedm ? "FFMA2" : "FFMA ", best, 2.0 * 16 * it * (double)grid * 256 / best / 1e9);
    }
    return 0;
}
