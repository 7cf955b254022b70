// What can HBM deliver for the lookups' access pattern?  Each thread reads ONE short run of consecutive floats at a
// pseudo-random offset of a 1 GiB buffer (no reuse, > L2), sums it and writes one float: the gather of the correlation
// lookup without anything else (16 registers, full occupancy: 2048 threads per SM keep the memory system as busy as it
// can be kept).  Cases: 10 floats at 4-byte alignment (a RAFT-Stereo tap run), 32 floats on a 128-byte line, 80 floats at
// 32-byte alignment (an IGEV geometry run, 16-byte loads), and a streaming read as the control.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/dram_random_probe.cu -o tools/build/dram_random_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

template <int RUN, int ALIGN, int VEC>
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ buf, uint64_t nfloats, float* __restrict__ out,
                                                    uint64_t nruns, uint64_t seed) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nruns) return;
    uint64_t off = mix(t + seed) % ((nfloats - RUN) / ALIGN) * ALIGN;
    const float* p = buf + off;
    float s = 0.f;
    if (VEC == 4) {
#pragma unroll
        for (int k = 0; k < RUN; k += 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p + k));
            s += v.x + v.y + v.z + v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < RUN; ++k) s += __ldg(p + k);
    }
    out[t] = s;
}

__global__ void __launch_bounds__(256) stream_kernel(const float4* __restrict__ buf, uint64_t n4, float* __restrict__ out) {
    float s = 0.f;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(buf + i);
        s += v.x + v.y + v.z + v.w;
    }
    if (s == 1.2345f) out[0] = s;
}

// The same runs fetched by the TMA unit: one cp.async.bulk (1-D, RUN*4 bytes, 16-byte aligned) per thread per round into
// shared memory, completion on one mbarrier per block; the thread then reads its run back from shared memory.
template <int RUN, int ALIGN>
__global__ void __launch_bounds__(256) bulk_kernel(const float* __restrict__ buf, uint64_t nfloats, float* __restrict__ out,
                                                  uint64_t nruns, uint64_t seed) {
    extern __shared__ __align__(128) uint8_t smem[];
    float* stage = reinterpret_cast<float*>(smem);                       // [256][RUN]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 256 * RUN * 4);
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    float s = 0.f;
    const uint64_t rounds = (nruns + (uint64_t)gridDim.x * 256 - 1) / ((uint64_t)gridDim.x * 256);
    for (uint64_t r = 0; r < rounds; ++r) {
        const uint64_t t = (r * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(256u * RUN * 4u) : "memory");
        __syncthreads();
        const uint64_t off = mix(t + seed) % ((nfloats - RUN) / ALIGN) * ALIGN;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + threadIdx.x * RUN);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(buf + off), "r"((uint32_t)(RUN * 4)), "r"(bar_a) : "memory");
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
        phase ^= 1u;
        s += stage[threadIdx.x * RUN] + stage[threadIdx.x * RUN + RUN - 1];
        __syncthreads();
    }
    out[(uint64_t)blockIdx.x * 256 + threadIdx.x] = s;
}

template <int RUN, int ALIGN>
static void run_bulk(const char* name, const float* buf, uint64_t nfloats, float* out, uint64_t nruns, int ctas_sm) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const size_t smem = 256 * RUN * 4 + 64;
    cudaFuncSetAttribute(bulk_kernel<RUN, ALIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = 148u * ctas_sm;
    for (int i = 0; i < 3; ++i) bulk_kernel<RUN, ALIGN><<<grid, 256, smem>>>(buf, nfloats, out, nruns, 1000 + i);
    cudaEventRecord(a);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) bulk_kernel<RUN, ALIGN><<<grid, 256, smem>>>(buf, nfloats, out, nruns, 77 * i);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double us = ms * 1e3 / reps;
    printf("%-34s %d CTA/SM runs %8llu  %8.1f us  %6.2f G runs/s  useful %7.1f GB/s\n", name, ctas_sm,
           (unsigned long long)nruns, us, nruns / us * 1e-3, nruns * (double)RUN * 4 / us * 1e-3);
}

template <int RUN, int ALIGN, int VEC>
static void run_case(const char* name, const float* buf, uint64_t nfloats, float* out, uint64_t nruns) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const unsigned grid = (unsigned)((nruns + 255) / 256);
    for (int i = 0; i < 3; ++i) gather_kernel<RUN, ALIGN, VEC><<<grid, 256>>>(buf, nfloats, out, nruns, 1000 + i);
    cudaEventRecord(a);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) gather_kernel<RUN, ALIGN, VEC><<<grid, 256>>>(buf, nfloats, out, nruns, 77 * i);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    const double us = ms * 1e3 / reps;
    // lines a run touches on average: (RUN*4 - ALIGN*4) / 128 + 1 for ALIGN*4 <= 128
    const double lines = ((double)RUN * 4 - ALIGN * 4) / 128.0 + 1.0;
    printf("%-44s runs %8llu  %8.1f us  %6.2f G runs/s  useful %7.1f GB/s  lines (128 B) %6.2f G/s = %7.1f GB/s\n", name,
           (unsigned long long)nruns, us, nruns / us * 1e-3, nruns * (double)RUN * 4 / us * 1e-3, nruns * lines / us * 1e-3,
           nruns * lines * 128 / us * 1e-3);
}

int main() {
    const uint64_t nfloats = 1ull << 28;      // 1 GiB
    float *buf, *out;
    cudaMalloc(&buf, nfloats * 4);
    cudaMalloc(&out, (64ull << 20));
    cudaMemset(buf, 0, nfloats * 4);
    for (uint64_t nruns : {1044480ull, 4177920ull}) {
        run_case<10, 1, 1>("10 floats, 4-byte aligned (RAFT tap run)", buf, nfloats, out, nruns);
        run_case<32, 32, 4>("32 floats on a 128-byte line (16-byte loads)", buf, nfloats, out, nruns);
        run_case<8, 8, 4>("8 floats on a 32-byte sector", buf, nfloats, out, nruns);
        run_case<80, 8, 4>("80 floats, 32-byte aligned (IGEV geo run)", buf, nfloats, out, nruns);
    }
    for (int ctas : {1, 2, 4, 8}) {
        if (ctas <= 2) run_bulk<80, 8>("TMA bulk 320 B, 32-byte aligned", buf, nfloats, out, 4177920ull, ctas);
        run_bulk<16, 4>("TMA bulk 64 B, 16-byte aligned", buf, nfloats, out, 4177920ull, ctas);
        run_bulk<32, 4>("TMA bulk 128 B, 16-byte aligned", buf, nfloats, out, 4177920ull, ctas);
    }
    {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        stream_kernel<<<148 * 8, 256>>>(reinterpret_cast<const float4*>(buf), nfloats / 4, out);
        cudaEventRecord(a);
        for (int i = 0; i < 5; ++i) stream_kernel<<<148 * 8, 256>>>(reinterpret_cast<const float4*>(buf), nfloats / 4, out);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        printf("streaming read of 1 GiB: %.1f us = %.1f GB/s\n", ms * 1e3 / 5, nfloats * 4.0 / (ms / 5 * 1e-3) * 1e-9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
