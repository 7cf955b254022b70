// Probe (B200): does a tcgen05.mma K-major SWIZZLE_128B shared-memory descriptor tolerate
//   (a) a start address that is not 1024-byte aligned (any multiple of 128 B), and
//   (b) a stride between 8-row atoms (SBO) that is not a multiple of 1024 B (here 1280 B = 10 rows)?
// If yes, one TMA halo patch of a 16 x 8-pixel tile ((16+2) x (8+2) pixels, 128 B per pixel) can serve all nine taps of
// a 3x3 convolution: tap (ky,kx) = the same patch read from row offset ky*10 + kx with atom stride 10 rows
// (DESIGN.md "what comes next", item 3).
//
// Method: A = a [ROWS][64] bf16 matrix whose row r holds the value r in column (r % 64) and 0 elsewhere... simpler:
// A[r][k] = (k == 0) ? r : 0 ; B = [64][64] with B[n][0] = 1 for all n.  Then D[m][n] = A[row(m)][0] = row(m): the
// accumulator directly shows which shared-memory row fed M-row m.  A is brought in by TMA (SWIZZLE_128B), exactly as in
// the conv kernels.  Expected row(m) = start_row + (m / 8) * atom_rows + (m % 8).
//
// Build / run (GPU box):  nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe tools/umma_stride_probe.cu && /tmp/probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../dkt_stereo_b200/csrc/tc.cuh"

using namespace dkt::tc;

constexpr int ROWS = 200;          // smem rows of A (>= 7*... 2*10+2 + 15*10 + 8)

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

struct Params {
    CUtensorMap a, b;
    float* out;            // [128][64]
    int start_row, atom_rows, use_base_off;
};

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ Params prm) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                          // ROWS x 128 B
    uint8_t* sb = smem + 32 * 1024;              // 64 x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(slot, 64);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar[0], ROWS * 128 + 64 * 128);
        tma_load_2d(sa, &prm.a, &bar[0], 0, 0);
        tma_load_2d(sb, &prm.b, &bar[0], 0, 0);
    }
    mbar_wait(&bar[0], 0);
    tcgen05_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(sa) + (uint32_t)prm.start_row * 128u;
        const uint32_t boff = prm.use_base_off ? ((a0 >> 7) & 7u) : 0u;
        const uint32_t idesc = idesc_bf16_m128(64);
        for (int k = 0; k < 4; ++k)
            umma_bf16(tmem, desc_sw128(a0 + k * 32, (uint32_t)prm.atom_rows * 128u, boff), desc_sw128(smem_u32(sb) + k * 32, 1024, 0), idesc, k != 0);
        umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tcgen05_fence_after();
    float v[32];
    for (int c = 0; c < 64; c += 32) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) prm.out[(warp * 32 + lane) * 64 + c + j] = v[j];
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
    std::vector<__nv_bfloat16> ha((size_t)ROWS * 64), hb(64 * 64);
    for (int r = 0; r < ROWS; ++r)
        for (int k = 0; k < 64; ++k) ha[(size_t)r * 64 + k] = __float2bfloat16(k == 5 ? (float)r : 0.f);   // row id in column 5
    for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) hb[n * 64 + k] = __float2bfloat16(k == 5 ? 1.f : 0.f);
    __nv_bfloat16 *da, *db;
    float* dout;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    Params prm{};
    const uint64_t dA[2] = {64, ROWS}, sA[2] = {1, 64};
    const uint32_t bA[2] = {64, ROWS};
    const uint64_t dB[2] = {64, 64}, sB[2] = {1, 64};
    const uint32_t bB[2] = {64, 64};
    if (!make_tmap_bf16(&prm.a, da, 2, dA, sA, bA) || !make_tmap_bf16(&prm.b, db, 2, dB, sB, bB)) {
        printf("tensor map encode failed\n");
        return 1;
    }
    prm.out = dout;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int cases[][3] = {{0, 8, 0}, {16, 8, 0}, {3, 8, 0}, {3, 8, 1}, {0, 10, 0}, {0, 10, 1}, {13, 10, 0}, {13, 10, 1}, {21, 10, 1}};
    std::vector<float> ho(128 * 64);
    for (auto& c : cases) {
        prm.start_row = c[0]; prm.atom_rows = c[1]; prm.use_base_off = c[2];
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe_kernel<<<1, 128, 64 * 1024>>>(prm);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("start %d atom_rows %d base_off %d: CUDA error %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first_bad = -1;
        for (int m = 0; m < 128; ++m) {
            const float want = (float)(c[0] + (m / 8) * c[1] + (m % 8));
            if (ho[m * 64 + 0] != want) { if (first_bad < 0) first_bad = m; ++bad; }
        }
        printf("start_row %2d  atom stride %2d rows  base_offset %s : %s", c[0], c[1], c[2] ? "set " : "zero", bad ? "MISMATCH" : "ok");
        if (bad) printf(" (%d rows, first m=%d got %.0f want %d)", bad, first_bad, ho[first_bad * 64], c[0] + (first_bad / 8) * c[1] + (first_bad % 8));
        printf("\n");
    }
    return 0;
}
