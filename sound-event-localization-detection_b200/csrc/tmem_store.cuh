// Tensor memory (TMEM, 256 KB per SM on sm_100a) used as a PER-LANE store: with the 32x32b access shape thread i of a
// warp reads / writes `n` consecutive 32-bit columns of TMEM lane 32 * (warp % 4) + i — the same thread always sees the
// same cells, nothing crosses lanes.  That is exactly the access pattern of
//   * the per-lane constant tables (window row, twiddle row of a lane) that every frame of the feature kernels re-reads and
//     that otherwise go through shared memory and its 128 B/clk datapath, the busiest pipe of the FOA kernel;
//   * the unit phasors a lane of the MIC kernel writes once per frame and reads back three times: out of
//     shared memory and registers they are what lets that kernel run 12 warps per SM instead of 8.
// No tensor-core instruction is involved: tcgen05.alloc / st / ld / dealloc only.
// Measured on B200: wide accesses to read-only tables fetched ahead of independent work pay (three LDTM.x32 for 24 LDS.128:
// -4 % FOA kernel time); narrow ones do not (x4 / x8: +14 %), and neither do loads on a dependent path that buy no
// occupancy (the FOA kernel's parked spectrum with x32 / x16 accesses: +8.7 %).
#pragma once
#include <cstdint>

namespace seld {
namespace tmem {

// Allocations are a power of two >= 32 columns (of the SM's 512); a kernel takes only what its tables need, so another
// kernel's CTA that shares the SM (none fits next to the feature kernels' shared memory today) still finds columns.
// warp-collective; `slot` is a shared-memory word that receives the base address
template <int COLS>
__device__ __forceinline__ void alloc(uint32_t* slot) {
    static_assert(COLS >= 32 && COLS <= 512 && (COLS & (COLS - 1)) == 0, "tensor memory columns");
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// address of column `col` in this warp's 32 lanes
__device__ __forceinline__ uint32_t lane_base(uint32_t base, int warp) { return base + ((uint32_t)(32 * (warp & 3)) << 16); }

// Stores complete at the next wait_st(); loads are issued and completed in ONE asm statement (ld + wait::ld), so the
// compiler can never use a destination register before the wait.
__device__ __forceinline__ void st16(uint32_t addr, const float (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(addr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "f"(r[8]), "f"(r[9]),
          "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]));
}
__device__ __forceinline__ void st32(uint32_t addr, const float (&r)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(addr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]), "f"(r[16]), "f"(r[17]), "f"(r[18]), "f"(r[19]), "f"(r[20]), "f"(r[21]), "f"(r[22]), "f"(r[23]), "f"(r[24]), "f"(r[25]), "f"(r[26]), "f"(r[27]), "f"(r[28]), "f"(r[29]), "f"(r[30]), "f"(r[31]));
}
__device__ __forceinline__ void ld32(uint32_t addr, float (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32]; tcgen05.wait::ld.sync.aligned;"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]),
          "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]),
          "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]),
          "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(addr));
}

// Split form: the load is issued here and completed by ldN_wait(r) — code in between overlaps the TMEM access.  The wait
// names every destination register as an in/out operand, so no use of r[] can be scheduled ahead of it.
__device__ __forceinline__ void ld32_issue(uint32_t addr, float (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]), "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]), "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
                 : "r"(addr));
}
__device__ __forceinline__ void ld32_wait(float (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]), "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]), "+f"(r[14]), "+f"(r[15]), "+f"(r[16]), "+f"(r[17]), "+f"(r[18]), "+f"(r[19]), "+f"(r[20]), "+f"(r[21]), "+f"(r[22]), "+f"(r[23]), "+f"(r[24]), "+f"(r[25]), "+f"(r[26]), "+f"(r[27]), "+f"(r[28]), "+f"(r[29]), "+f"(r[30]), "+f"(r[31]));
}

}  // namespace tmem
}  // namespace seld
