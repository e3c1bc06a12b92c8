// smem_probe.cu — measures shared-memory wavefronts per instruction for chosen lane->address patterns
// (LDS.128 / STS.128 / LDS.32) with one warp per CTA, by clock64 over a dependent-free loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/smem_probe.bin tools/smem_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 v;
    unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(float4* p, float4 v) {
    unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    asm volatile("st.shared.v4.f32 [%4], {%0,%1,%2,%3};" ::"f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(a) : "memory");
}
__device__ __forceinline__ float lds32(const float* p) {
    float v;
    unsigned a = static_cast<unsigned>(__cvta_generic_to_shared(p));
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}

template <int MODE>   // 0 = LDS.128, 1 = STS.128, 2 = LDS.32, 3 = RMW.128 (LDS + STS)
__global__ void probe(const int* __restrict__ idx, const int* __restrict__ pred, long long* out, float* sink) {
    __shared__ float4 sm[2048];
    const int lane = threadIdx.x;
    for (int k = lane; k < 2048; k += 32) sm[k] = make_float4(k, 0, 0, 0);
    __syncwarp();
    const int i = idx[lane];
    const bool on = pred[lane] != 0;
    float4 acc = make_float4(0, 0, 0, 0);
    float* smf = reinterpret_cast<float*>(sm);
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {
            if (on) {
                float4 v = lds128(&sm[i]);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        } else if (MODE == 1) {
            if (on) sts128(&sm[i], acc);
        } else if (MODE == 2) {
            if (on) acc.x += lds32(&smf[i]);
        } else {
            float4 v = lds128(&sm[i]);
            v.x += 1.0f;
            if (on) sts128(&sm[i], v);
        }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 32 + lane] = acc.x + acc.y + acc.z + acc.w;
}

static void run(const char* name, int mode, const std::vector<int>& idx, const std::vector<int>& pred) {
    int *d_idx, *d_pred; long long* d_out; float* d_sink;
    cudaMalloc(&d_idx, 32 * 4); cudaMalloc(&d_pred, 32 * 4); cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 32 * 4);
    cudaMemcpy(d_idx, idx.data(), 32 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_pred, pred.data(), 32 * 4, cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) probe<0><<<1, 32>>>(d_idx, d_pred, d_out, d_sink);
        if (mode == 1) probe<1><<<1, 32>>>(d_idx, d_pred, d_out, d_sink);
        if (mode == 2) probe<2><<<1, 32>>>(d_idx, d_pred, d_out, d_sink);
        if (mode == 3) probe<3><<<1, 32>>>(d_idx, d_pred, d_out, d_sink);
        cudaDeviceSynchronize();
    }
    long long t; cudaMemcpy(&t, d_out, 8, cudaMemcpyDeviceToHost);
    const char* mn[] = {"LDS.128", "STS.128", "LDS.32", "RMW.128"};
    printf("%-8s %-44s %7.2f cycles/iter\n", mn[mode], name, double(t) / ITERS);
    cudaFree(d_idx); cudaFree(d_pred); cudaFree(d_out); cudaFree(d_sink);
}

int main() {
    std::vector<int> all(32, 1), idx(32);
    auto P = [&](const char* name, auto f, std::vector<int> pred) {
        for (int l = 0; l < 32; ++l) idx[l] = f(l);
        for (int m : {0, 1, 3}) run(name, m, idx, pred);
    };
    P("consecutive", [](int l) { return l; }, all);
    P("all lanes one group (8l)", [](int l) { return 8 * l; }, all);
    P("quarters distinct, quarters collide", [](int l) { return (l % 8) + 40 * (l / 8); }, all);
    P("quarter rows shifted (l%8 + 8*3*(l/8))", [](int l) { return (l % 8) + 24 * (l / 8); }, all);
    P("one pair in quarter 0 collides", [](int l) { return l == 1 ? 8 : (l == 0 ? 0 : l + 64); }, all);
    P("one pair per quarter collides", [](int l) { return (l % 8 == 1) ? 200 + 8 * (l / 8) * 8 : ((l % 8 == 0) ? 8 * (l / 8) * 8 : l + 512); }, all);
    P("lanes l, l+4 same group (half-quarter)", [](int l) { return (l % 4) + 8 * (l / 4) * 3; }, all);
    P("groups distinct per quarter, scrambled", [](int l) { int q = l / 8, r = l % 8; return ((r * 3 + q) % 8) + 8 * (5 * q + r); }, all);
    P("lanes 0..15 distinct per HALF only (16 groups?)", [](int l) { return (l % 16) + 16 * 7 * (l / 16); }, all);
    P("same address all lanes", [](int l) { return 5; }, all);
    P("same address per quarter", [](int l) { return 8 * (l / 8); }, all);
    {
        std::vector<int> half(32, 0);
        for (int l = 0; l < 16; ++l) half[l] = 1;
        P("consecutive, lanes 16..31 off", [](int l) { return l; }, half);
        std::vector<int> q3(32, 1);
        for (int l = 24; l < 32; ++l) q3[l] = 0;
        P("consecutive, quarter 3 off", [](int l) { return l; }, q3);
        std::vector<int> odd(32, 0);
        for (int l = 0; l < 32; l += 2) odd[l] = 1;
        P("consecutive, odd lanes off", [](int l) { return l; }, odd);
        P("even lanes: 16 distinct groups across warp", [](int l) { return l / 2 + 8 * 9 * (l % 2); }, odd);
        // idle quarters whose lanes mirror lane 0's address (what the tile kernels do): is the quarter free?
        P("quarter 3 off, mirrors lane 0", [](int l) { return l < 24 ? l : 0; }, q3);
        P("lanes 16..31 off, mirror lane 0", [](int l) { return l < 16 ? l : 0; }, half);
        std::vector<int> q0(32, 0);
        for (int l = 0; l < 8; ++l) q0[l] = 1;
        P("only quarter 0 on, rest mirror lane 0", [](int l) { return l < 8 ? l : 0; }, q0);
        P("only quarter 0 on, rest consecutive", [](int l) { return l; }, q0);
        std::vector<int> l20(32, 0);
        for (int l = 0; l < 20; ++l) l20[l] = 1;
        P("lanes 0..19 on, 20..23 mirror lane 16, q3 mirrors 0", [](int l) { return l < 20 ? l : (l < 24 ? 16 : 0); }, l20);
    }
    // scalar
    auto S = [&](const char* name, auto f) {
        for (int l = 0; l < 32; ++l) idx[l] = f(l);
        run(name, 2, idx, all);
    };
    S("scalar consecutive", [](int l) { return l; });
    S("scalar 2-way (l/2*... )", [](int l) { return (l % 16) + 32 * (l / 16) * 3; });
    S("scalar 4-way", [](int l) { return (l % 8) + 32 * (l / 8) * 3; });
    S("scalar stride 12 rows of 10 (x+12y)", [](int l) { return (l % 10) + 12 * (l / 10); });
    return 0;
}
