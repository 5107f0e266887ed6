// Block-level staging of the optical-table window and the layer list into shared memory.
//
// sm_100a path: the bulk-copy engine (TMA, 1-D form: cp.async.bulk.shared::cluster.global with an mbarrier
// transaction count) moves the tables; one elected thread programs the copies, the block waits on the mbarrier
// phase.  The table layout in HBM was chosen for this: a block's wavelength window of `opt` ([wl][mat][6], 48-byte
// rows) is one contiguous 16-byte-aligned range, and the layer list is one contiguous array of 16-byte entries.
// Rows are re-pitched in shared memory (odd multiple of 16 B) so that lanes on different wavelengths hit different
// bank groups; with a pitch equal to the row size the whole window is a single copy, otherwise one copy per row.
// Fallback (PST_FLAG_NO_BULK_COPY, or a range that is not 16-byte aligned): cooperative 128-bit loads.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pst {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// Wait for phase `parity` of the barrier.  The spin is bounded: a mis-programmed transaction count traps instead of
// hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    if (spins > (1u << 22)) __trap();
  }
}

// Stage `rows` rows of `row_bytes` bytes (contiguous in global memory starting at `src`) into shared memory with
// row pitch `pitch`, plus `extra_bytes` from `extra_src` to `extra_dst` (the layer list).  All sizes are multiples of
// 16 and all addresses 16-byte aligned when `use_bulk` is set by the host.  Ends with a block-wide barrier.
__device__ __forceinline__ void stage_block(bool use_bulk, unsigned char* dst, const unsigned char* src, int rows,
                                            int row_bytes, int pitch, unsigned char* extra_dst,
                                            const unsigned char* extra_src, int extra_bytes, int tid, int nthr) {
  __shared__ __align__(8) uint64_t s_bar;
  if (use_bulk) {
    if (tid == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid < 32) {
      const uint32_t total = (uint32_t)rows * (uint32_t)row_bytes + (uint32_t)extra_bytes;
      if (tid == 0) mbar_expect_tx(&s_bar, total);
      __syncwarp();
      if (rows > 0) {
        if (pitch == row_bytes) {
          if (tid == 0) bulk_g2s(dst, src, (uint32_t)rows * (uint32_t)row_bytes, &s_bar);
        } else {
          for (int r = tid; r < rows; r += 32)
            bulk_g2s(dst + (size_t)r * pitch, src + (size_t)r * row_bytes, (uint32_t)row_bytes, &s_bar);
        }
      }
      if (extra_bytes > 0 && tid == 0) bulk_g2s(extra_dst, extra_src, (uint32_t)extra_bytes, &s_bar);
    }
    mbar_wait(&s_bar, 0);
    return;
  }
  const int per_row = row_bytes / 16;
  const int4* s4 = reinterpret_cast<const int4*>(src);
  for (int idx = tid; idx < per_row * rows; idx += nthr) {
    const int r = idx / per_row, c = idx - r * per_row;
    *reinterpret_cast<int4*>(dst + (size_t)r * pitch + c * 16) = __ldg(s4 + idx);
  }
  const int4* e4 = reinterpret_cast<const int4*>(extra_src);
  for (int idx = tid; idx < extra_bytes / 16; idx += nthr) reinterpret_cast<int4*>(extra_dst)[idx] = __ldg(e4 + idx);
  __syncthreads();
}

}  // namespace pst
