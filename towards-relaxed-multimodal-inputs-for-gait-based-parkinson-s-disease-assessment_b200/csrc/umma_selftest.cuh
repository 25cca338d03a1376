// umma_selftest.cuh -- hardware self-test of the tcgen05 layer: runs an arbitrary list of tf32 MMAs over
// caller-supplied shared-memory images of A and B and returns the 128 x N fp32 accumulator.  Used by
// tests/test_gpu_umma.py to pin the descriptor conventions (K-major tap-shifted operands for the
// convolutions, MN-major operands for the weight gradients) against a CPU product.
#pragma once
#include "umma.cuh"

namespace gaitk {

struct UmmaOp { uint32_t a_off, a_lbo, a_sbo, b_off, b_lbo, b_sbo, accumulate, idesc; };

__global__ void __launch_bounds__(128) umma_selftest_kernel(const float* A, int nA, const float* B, int nB, const UmmaOp* ops,
                                                            int nops, int ncols, float* D /*128 x ncols*/) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float* As = sm; float* Bs = sm + ((nA + 255) / 256) * 256;
    const int tid = threadIdx.x;
    for (int i = tid; i < nA; i += 128) As[i] = umma::to_tf32(A[i]);
    for (int i = tid; i < nB; i += 128) Bs[i] = umma::to_tf32(B[i]);
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    if (tid < 32) umma::tmem_alloc(&tmem_slot, (uint32_t)ncols);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        const uint32_t a0 = umma::smem_u32(As), b0 = umma::smem_u32(Bs);
        for (int i = 0; i < nops; ++i) {
            const UmmaOp o = ops[i];
            umma::mma_tf32(tbase, umma::make_desc(a0 + o.a_off, o.a_lbo, o.a_sbo), umma::make_desc(b0 + o.b_off, o.b_lbo, o.b_sbo),
                           o.idesc, o.accumulate);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    const uint32_t lane_base = (uint32_t)(tid & ~31) << 16;
    for (int c = 0; c < ncols; c += 8) {
        float v[8];
        umma::ld_x8(tbase + lane_base + c, v);
        umma::ld_wait();
        for (int j = 0; j < 8; ++j) D[(size_t)tid * ncols + c + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tbase, (uint32_t)ncols);
}

// Same harness for kind::f16 / bf16 operands: A and B are raw 16-bit images (already bf16 bit patterns, passed as
// uint16), copied verbatim; pins the K-major and MN-major no-swizzle conventions of the split-bf16 stream kernel.
__global__ void __launch_bounds__(128) umma_selftest_bf16_kernel(const uint16_t* A, int nA, const uint16_t* B, int nB, const UmmaOp* ops,
                                                                 int nops, int ncols, float* D /*128 x ncols*/) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint16_t* As = reinterpret_cast<uint16_t*>(sm); uint16_t* Bs = As + ((nA + 511) / 512) * 512;
    const int tid = threadIdx.x;
    for (int i = tid; i < nA; i += 128) As[i] = A[i];
    for (int i = tid; i < nB; i += 128) Bs[i] = B[i];
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    if (tid < 32) umma::tmem_alloc(&tmem_slot, (uint32_t)ncols);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        const uint32_t a0 = umma::smem_u32(As), b0 = umma::smem_u32(Bs);
        for (int i = 0; i < nops; ++i) {
            const UmmaOp o = ops[i];
            // bit 31 of `accumulate` = column offset of D in units of 8 columns (bits 8..15)
            const uint32_t dcol = (o.accumulate >> 8) & 0xffu;
            umma::mma_bf16(tbase + dcol * 8u, umma::make_desc(a0 + o.a_off, o.a_lbo, o.a_sbo), umma::make_desc(b0 + o.b_off, o.b_lbo, o.b_sbo),
                           o.idesc, o.accumulate & 1u);
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    const uint32_t lane_base = (uint32_t)(tid & ~31) << 16;
    for (int c = 0; c < ncols; c += 8) {
        float v[8];
        umma::ld_x8(tbase + lane_base + c, v);
        umma::ld_wait();
        for (int j = 0; j < 8; ++j) D[(size_t)tid * ncols + c + j] = v[j];
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tbase, (uint32_t)ncols);
}

// Cost of a tcgen05.mma by shape / operand layout: one thread issues the op list `reps` times back to back (all ops
// accumulate), commits once and waits; cycles[0] = SM clocks from the first issue to completion, cycles[1] = clocks spent
// issuing.  Operands are whatever shared memory holds (zeros): only time matters.
__global__ void __launch_bounds__(128) umma_bench_kernel(const UmmaOp* ops, int nops, int reps, int ncols, int smem_bytes, long long* cycles) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x;
    for (int i = tid; i < smem_bytes / 4; i += 128) sm[i] = 0.f;
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
    if (tid < 32) umma::tmem_alloc(&tmem_slot, (uint32_t)ncols);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        const uint32_t a0 = umma::smem_u32(sm);
        // descriptors in registers (up to 4 distinct ops, cycled): the loop measures the tensor pipe, not operand loads
        uint64_t da[4], db[4]; uint32_t dd[4], id[4];
        for (int i = 0; i < 4; ++i) {
            const UmmaOp o = ops[i % nops];
            da[i] = umma::make_desc(a0 + o.a_off, o.a_lbo, o.a_sbo); db[i] = umma::make_desc(a0 + o.b_off, o.b_lbo, o.b_sbo);
            dd[i] = tbase + ((o.accumulate >> 8) & 0xffu) * 8u; id[i] = o.idesc;
        }
        const long long t0 = clock64();
        for (int r = 0; r < reps * nops; r += 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) umma::mma_bf16(dd[i], da[i], db[i], id[i], 1u);
        }
        const long long t1 = clock64();
        umma::commit(&bar);
        umma::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        cycles[0] = t2 - t0; cycles[1] = t1 - t0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tbase, (uint32_t)ncols);
}

// The same with `nissue` warps issuing concurrently (warp w -> its own mbarrier, D columns shifted by 64 w): does the ~45-clock
// cost of a small tcgen05.mma belong to the issuing THREAD (then n issuers give n times the rate) or to the tensor pipe
// (then the rate stays)?  cycles[2 w] = issuer w's clocks first issue -> completion, cycles[2 w + 1] = clocks spent issuing.
__global__ void __launch_bounds__(128) umma_bench_multi_kernel(const UmmaOp* ops, int nops, int reps, int ncols, int smem_bytes, int nissue, long long* cycles) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < smem_bytes / 4; i += 128) sm[i] = 0.f;
    if (tid == 0) { for (int i = 0; i < 4; ++i) umma::mbar_init(&bar[i], 1); umma::fence_mbar_init(); }
    if (tid < 32) umma::tmem_alloc(&tmem_slot, (uint32_t)ncols);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if ((tid & 31) == 0 && warp < nissue) {
        const uint32_t a0 = umma::smem_u32(sm);
        uint64_t da[4], db[4]; uint32_t dd[4], id[4];
        for (int i = 0; i < 4; ++i) {
            const UmmaOp o = ops[i % nops];
            da[i] = umma::make_desc(a0 + o.a_off, o.a_lbo, o.a_sbo); db[i] = umma::make_desc(a0 + o.b_off, o.b_lbo, o.b_sbo);
            dd[i] = tbase + ((o.accumulate >> 8) & 0xffu) * 8u + 64u * warp; id[i] = o.idesc;
        }
        const long long t0 = clock64();
        for (int r = 0; r < reps * nops; r += 4) {
#pragma unroll
            for (int i = 0; i < 4; ++i) umma::mma_bf16(dd[i], da[i], db[i], id[i], 1u);
        }
        const long long t1 = clock64();
        umma::commit(&bar[warp]);
        umma::mbar_wait(&bar[warp], 0);
        const long long t2 = clock64();
        cycles[2 * warp] = t2 - t0; cycles[2 * warp + 1] = t1 - t0;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (tid < 32) umma::tmem_dealloc(tbase, (uint32_t)ncols);
}

}  // namespace gaitk
