// Development test (GPU): A = J^T J on the 5th-generation tensor cores with the 3xTF32 split (hi*hi + lo*hi + hi*lo),
// operands written by ordinary threads into the canonical no-swizzle K-major shared-memory layout, accumulator in
// tensor memory.  Validates the descriptor encodings used by the Stage-II kernel before they go into it.
//   built by moshpp_b200.build.build_tc_test(), run by tests/test_gpu_parity.py::test_tcgen05_jtj_building_block
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

namespace tc {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ float to_tf32(float v) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return __uint_as_float(r); }
// canonical K-major, no swizzle: core matrix = 8 rows x 16 bytes; LBO = distance between the core matrices along K,
// SBO = distance between 8-row groups (both in bytes)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= uint64_t((addr & 0x3FFFF) >> 4);
    d |= uint64_t((lbo >> 4) & 0x3FFF) << 16;
    d |= uint64_t((sbo >> 4) & 0x3FFF) << 32;
    d |= uint64_t(1) << 46;                 // descriptor version (Blackwell)
    return d;                               // layout type 0: no swizzle
}
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
}  // namespace tc

constexpr int MM = 128;         // UMMA M (free variables, zero padded)

// X[i][k] (i: column of J, k: row of J within the tile) -> float index inside a [MM][KT] canonical buffer
__device__ __host__ inline int xidx(int i, int k, int KT) {
    // core matrices along K are adjacent (LBO = 128 B); 8-row groups are KT/4 core matrices apart (SBO)
    return (i >> 3) * (KT / 4) * 32 + (k >> 2) * 32 + (i & 7) * 4 + (k & 3);
}

__global__ void __launch_bounds__(384, 1) jtj_kernel(const float *J, int rows, int n, float *D, int mode, int KT, int skew) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *Xhi = reinterpret_cast<float *>(smem_raw + skew);      // skew: operand base not 128-byte aligned
    float *Xlo = Xhi + MM * KT;
    float *Xl2 = Xlo + MM * KT;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + 3 * MM * KT * 4 + 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = (n + 15) & ~15;
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(tmem_slot), 512);
    if (tid == 32) tc::mbar_init(tc::smem_u32(mbar), 1);
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = *tmem_slot;
    uint32_t phase = 0;
    const uint32_t idesc = tc::idesc_tf32(MM, N);
    int tiles = 0, kstep = 0;
    for (int r0 = 0; r0 < rows; r0 += KT, ++tiles) {
        for (int e = tid; e < MM * KT; e += blockDim.x) {
            const int i = e / KT, k = e - i * KT;
            float v = 0.f;
            if (i < n && r0 + k < rows) v = J[(r0 + k) * n + i];
            const float hi = tc::to_tf32(v), lo = tc::to_tf32(v - hi), l2 = (v - hi) - lo;
            Xhi[xidx(i, k, KT)] = hi;
            Xlo[xidx(i, k, KT)] = lo;
            Xl2[xidx(i, k, KT)] = l2;
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after();
            const uint32_t ahi = tc::smem_u32(Xhi), alo = tc::smem_u32(Xlo), al2 = tc::smem_u32(Xl2);
            const uint32_t LBO = 128, SBO = (KT / 4) * 128;
            for (int ks = 0; ks < KT / 8; ++ks) {
                const uint64_t dhi = tc::smem_desc(ahi + ks * 2 * LBO, LBO, SBO), dlo = tc::smem_desc(alo + ks * 2 * LBO, LBO, SBO);
                if (mode <= 3) {
                    tc::mma_tf32(tmem, dhi, dhi, idesc, (tiles > 0 || ks > 0) ? 1u : 0u);
                    if (mode == 3) {
                        tc::mma_tf32(tmem, dlo, dhi, idesc, 1u);
                        tc::mma_tf32(tmem, dhi, dlo, idesc, 1u);
                    }
                } else {
                    // mode 5 (the Stage-II kernel's scheme), 6, 7: hi*hi alternates between two (mode 7: three) accumulators,
                    // the small cross terms go to their own
                    const uint64_t dl2 = tc::smem_desc(al2 + ks * 2 * LBO, LBO, SBO);
                    const int nacc = mode == 7 ? 3 : 2, which = kstep % nacc;
                    tc::mma_tf32(tmem + 128 * which, dhi, dhi, idesc, kstep >= nacc ? 1u : 0u);
                    tc::mma_tf32(tmem + 384, dlo, dhi, idesc, kstep > 0 ? 1u : 0u);
                    tc::mma_tf32(tmem + 384, dhi, dlo, idesc, 1u);
                    tc::mma_tf32(tmem + 384, dlo, dlo, idesc, 1u);
                    if (mode >= 6) {                                // 6, 7: a third split level (not used by the Stage-II kernel)
                        tc::mma_tf32(tmem + 384, dl2, dhi, idesc, 1u);
                        tc::mma_tf32(tmem + 384, dhi, dl2, idesc, 1u);
                    }
                    ++kstep;
                }
            }
            tc::commit(tc::smem_u32(mbar));
        }
        tc::mbar_wait(tc::smem_u32(mbar), phase);     // the tile buffers are free again (and, at the end, D is complete)
        phase ^= 1;
    }
    tc::fence_after();
    if (warp < 4) {
        const int i = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
            if (mode >= 5) {
                float v1[16], v2[16], v3[16];
                tc::tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + 128 + c0, v1);
                tc::tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + 384 + c0, v3);
                if (mode == 7) tc::tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + 256 + c0, v2);
                for (int q = 0; q < 16; ++q) v[q] = ((v[q] + v1[q]) + (mode == 7 ? v2[q] : 0.f)) + v3[q];
            }
            if (i < n) for (int q = 0; q < 16; ++q) if (c0 + q < n) D[i * n + c0 + q] = v[q];
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

static int run(int rows, int n, int KT, int skew) {
    std::vector<float> J(rows * n);
    srand(1);
    for (auto &v : J) v = (rand() / float(RAND_MAX) - 0.5f) * ((rand() % 7 == 0) ? 40.f : 1.f);
    std::vector<double> ref(n * n, 0.0);
    for (int r = 0; r < rows; ++r)
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) ref[i * n + j] += double(J[r * n + i]) * double(J[r * n + j]);
    float *dJ, *dD;
    CK(cudaMalloc(&dJ, J.size() * 4)); CK(cudaMalloc(&dD, n * n * 4));
    CK(cudaMemcpy(dJ, J.data(), J.size() * 4, cudaMemcpyHostToDevice));
    const size_t smem = 3 * MM * KT * 4 + 256;
    CK(cudaFuncSetAttribute(jtj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    double scale = 0;
    for (int i = 0; i < n; ++i) scale = fmax(scale, ref[i * n + i]);
    int bad = 0;
    for (int mode : {3, 5, 6, 7}) {
        CK(cudaMemset(dD, 0, n * n * 4));
        jtj_kernel<<<1, 384, smem>>>(dJ, rows, n, dD, mode, KT, skew);
        CK(cudaDeviceSynchronize());
        std::vector<float> D(n * n);
        CK(cudaMemcpy(D.data(), dD, n * n * 4, cudaMemcpyDeviceToHost));
        double e = 0;
        for (int i = 0; i < n * n; ++i) e = fmax(e, fabs(D[i] - ref[i]));
        printf("rows %3d n %3d KT %2d skew %3d  %dxTF32: max |D - ref| / max diag = %.3e\n", rows, n, KT, skew, mode, e / scale);
        if (!(e / scale < (mode == 3 ? 1e-5 : 1.5e-6))) bad = 1;
    }
    cudaFree(dJ); cudaFree(dD);
    return bad;
}

int main() {
    int bad = 0;
    bad |= run(159, 111, 48, 0);
    bad |= run(159, 111, 64, 0);
    bad |= run(80, 30, 80, 0);
    bad |= run(80, 30, 80, 48);
    bad |= run(80, 6, 80, 48);
    bad |= run(40, 30, 40, 48);
    bad |= run(212, 111, 80, 0);
    printf(bad ? "FAILED\n" : "OK\n");
    return bad;
}
