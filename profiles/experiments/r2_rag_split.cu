// R1 raster pass, warp-specialised: pairs of a SCAN warp and an ITEM warp in one persistent CTA per SM.
//
// Same arithmetic as rag_blocks.cu (rag_core.cuh: 4 x 4 blocks, fast path for one-label windows, process_item for the
// rest); what changes is who runs it.  In the mixed kernel every warp alternates between the scan loop and item passes:
// 166 registers and 19 KB of shared memory per warp allow 12 warps per SM, and the kernel is bound by per-warp stalls
// (fixed-latency dependencies, branch resolution) at 3 warps per scheduler: 46 % of the issue slots.  Here
//   * the scan warp of a pair walks the strips: TMA-fed private stage ring, window test, unmasked statistics into the
//     lane's accumulators, a private region table for label changes.  A block that is not a one-label window is
//     written, as a self-contained item, straight into the pair's RING of item batches;
//   * the item warp of the pair owns the hash tables of the strip and does nothing but take full batches of 32 items
//     out of the ring, one item per lane (process_item).  (A CTA-wide ring feeding any item warp was measured first:
//     it loses the strip locality of the tables -- 2.7 x the raw edge entries, 0.76 ms.)
// Neither role needs more than ~95 registers once nothing is called out of line, so 18 warps fit where 12 did.
// The ring is NB batches of 32 slots; batch b has a FULL mbarrier (32 arrivals: one per item written, release) and a
// CONSUMED generation counter (stored, release, by the item warp after its pass; a scan lane about to write generation
// t of a batch waits until t generations have been consumed).  The scan warp numbers its items itself (one producer per
// ring: no atomics).  End of input: the scan warp publishes its item total, completes the last (partial) batch's
// barrier with the missing arrival count and raises `finished`; the item warp leaves when its next batch lies beyond
// the total.
#include "rag_tables.cuh"

namespace dm {
namespace rag {
namespace split {

using namespace blk;

template <int C_, int NP_, int NB_>
struct Cfg {
    static constexpr int C = C_, NP = NP_, NB = NB_;          // bands, warp pairs per CTA, batches per ring
    static constexpr int NSW = NP;
    static constexpr int TH = 4, NS = 2;
    static constexpr int CW = C_ > 0 ? C_ : 1;
    static constexpr int NWARPS = 2 * NP;
    static constexpr int THREADS = NWARPS * 32;
    static constexpr int LAB_BOX = align128(TH * LAB_PITCH * 4);
    static constexpr int IMG_ROW_WORDS = STRIP_W * C / 4;
    static constexpr int IMG_BOX = align128(TH * IMG_ROW_WORDS * 4);
    static constexpr int STAGE_BYTES = LAB_BOX + IMG_BOX;
    static constexpr int TX_BYTES = TH * LAB_PITCH * 4 + (C > 0 ? TH * IMG_ROW_WORDS * 4 : 0);
    static constexpr int ITEM_VECS = 7 + C;                  // above | 4 own rows | right column | C image vectors | position
    static constexpr int SLOTS = 32 * NB;
    static constexpr int RING_BYTES = SLOTS * ITEM_VECS * 16;
    static constexpr int TABLE_WORDS = RS * (3 + 2 * C) + ES * 3 + 4;
    static constexpr int TABLE_BYTES = align128(TABLE_WORDS * 4 + NS * 8);     // + the stage barriers of a scan warp
    static constexpr int CTRL_BYTES = 128;                                      // full[NB], consumed[NB], total, finished
    static constexpr int PAIR_BYTES = NS * STAGE_BYTES + 2 * TABLE_BYTES + RING_BYTES + CTRL_BYTES;
    static constexpr int SMEM_BYTES = 128 + NP * PAIR_BYTES;
    static constexpr int FLUSH_UNITS = FLUSH_ROWS / TH;
    static_assert(NB * 12 + 8 <= CTRL_BYTES, "control block");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct Ctrl {
    unsigned final_total;   // number of items of the pair, once known (~0u before)
    unsigned finished;      // final_total is valid and the last batch has been completed
};

__device__ __forceinline__ void mbar_arrive_release(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
// bounded wait: a barrier that never completes becomes an error code, not a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, unsigned parity, unsigned long long* counters) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 20)) {
            atomicExch(&counters[3], 2ull);
            __trap();
        }
    }
}

__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void wait_consumed(const unsigned* p, unsigned gen, unsigned long long* counters) {
    unsigned spins = 0;
    while (ld_acquire_u32(p) < gen) {
        __nanosleep(64);
        if (++spins > (1u << 22)) {
            atomicExch(&counters[3], 3ull);
            __trap();
        }
    }
}

template <int ICAP>
struct RingPick {
    const int* words;
    int j;
    __device__ __forceinline__ int operator()(int p) const {
        const int r = (p * 13) >> 6;             // p / 5 for p < 24
        const int k = p - 5 * r;
        const int vec = p >= 20 ? 0 : (k == 4 ? 5 : 1 + r);
        const int w = p >= 20 ? p - 20 : (k == 4 ? r : k);
        return words[(vec * ICAP + j) * 4 + w];
    }
};

template <typename CF, bool USE_TMA>
__global__ void __launch_bounds__(CF::THREADS, 1)
rag_split_kernel(const __grid_constant__ CUtensorMap mapL, const __grid_constant__ CUtensorMap mapI,
                 const __grid_constant__ Params P) {
    constexpr int C = CF::C, CW = CF::CW, TH = CF::TH, NS = CF::NS, NB = CF::NB, SLOTS = CF::SLOTS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    // Roles by scheduler: warps w with (w & 3) < 2 scan, the others process items, so that each of the four schedulers (and
    // its small L0 instruction cache) runs one role's code only (NP even); otherwise the first NP warps scan.
    constexpr bool BY_SMSP = (CF::NP % 2 == 0);
    const bool scan_role = BY_SMSP ? ((warp & 3) < 2) : (warp < CF::NP);
    const int pair = BY_SMSP ? ((warp >> 2) * 2 + (warp & 1)) : (warp < CF::NP ? warp : warp - CF::NP);
    unsigned char* pbase = smem + (size_t)pair * CF::PAIR_BYTES;       // this pair's arena
    unsigned char* wbase = pbase;                                      // stages | scan table | item table | ring | control
    uint4* ring = (uint4*)(pbase + NS * CF::STAGE_BYTES + 2 * CF::TABLE_BYTES);
    uint64_t* full_b = (uint64_t*)(pbase + NS * CF::STAGE_BYTES + 2 * CF::TABLE_BYTES + CF::RING_BYTES);   // [NB]
    unsigned* consumed = (unsigned*)(full_b + NB);                     // [NB] generations consumed of every batch
    Ctrl* ctrl = (Ctrl*)(consumed + NB);

    if (scan_role && lane == 0) {
        for (int b = 0; b < NB; ++b) {
            mbar_init(&full_b[b], 32);
            consumed[b] = 0;
        }
        ctrl->final_total = ~0u;
        ctrl->finished = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (scan_role) {
        // =============================================================================================================
        // SCAN role
        // =============================================================================================================
        unsigned* tab = (unsigned*)(wbase + NS * CF::STAGE_BYTES);
        const Tab<C> T = Tab<C>::from(tab);
        uint64_t* stage_bar = (uint64_t*)(tab + ((CF::TABLE_WORDS + 1) & ~1));
        unsigned produced = 0;                                             // items written to the ring so far (warp-uniform)

        const long long total_units = (long long)P.tiles_x * P.tiles_y;
        const long long gw = (long long)blockIdx.x * CF::NSW + pair;
        const long long u_begin = min(total_units, gw * (long long)P.tiles_per_cta);
        const long long u_end = min(total_units, u_begin + P.tiles_per_cta);
        const int my_units = (int)(u_end - u_begin);

        T.rkey[lane] = EMPTY_LABEL;
        T.rarea[lane] = 0;
        T.rborder[lane] = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            T.rsum[c * RS + lane] = 0;
            T.rsq[c * RS + lane] = 0;
        }
        T.ekey[lane] = EMPTY_KEY;
        T.ekey[lane + 32] = EMPTY_KEY;
        T.ecnt[lane] = 0;
        T.ecnt[lane + 32] = 0;
        if (lane == 0) {
            T.used[0] = T.used[1] = 0;
            if (USE_TMA) {
                for (int s = 0; s < NS; ++s) mbar_init(&stage_bar[s], 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
        }
        __syncwarp();

        int isx = (int)(u_begin / P.tiles_y), ij = (int)(u_begin - (long long)isx * P.tiles_y), issued = 0;
        auto issue = [&]() {
            if (lane == 0) {
                const int st = issued % NS;
                unsigned char* sb = wbase + (size_t)st * CF::STAGE_BYTES;
                mbar_expect_tx(&stage_bar[st], (unsigned)CF::TX_BYTES);
                tma_load_2d(sb, &mapL, isx * STRIP_W, ij * TH, &stage_bar[st]);
                if (C > 0) tma_load_2d(sb + CF::LAB_BOX, &mapI, isx * STRIP_W * C / 4, ij * TH, &stage_bar[st]);
            }
            ++issued;
            if (ij + 1 < P.tiles_y) ++ij;
            else { ij = 0; ++isx; }
        };
        if (USE_TMA) {
            for (int k = 0; k < NS && k < my_units; ++k) issue();
        }

        int4 up = make_int4(0, 0, 0, 0);
        int cur = EMPTY_LABEL;
        unsigned area = 0, s[CW], q[CW];
#pragma unroll
        for (int c = 0; c < CW; ++c) s[c] = q[c] = 0;
        const InlineSink<C> sink{T, P};
        auto lane_flush = [&]() {
            if (cur >= 0 && area) sink.region(cur, area, 0u, s, q);
            area = 0;
#pragma unroll
            for (int c = 0; c < CW; ++c) s[c] = q[c] = 0;
        };

        int units_since_flush = 0;
        int sx = (int)(u_begin / P.tiles_y), j = (int)(u_begin - (long long)sx * P.tiles_y);
        bool contiguous = false;

        for (int i = 0; i < my_units; ++i) {
            const int st = USE_TMA ? i % NS : 0;
            unsigned char* sb = wbase + (size_t)st * CF::STAGE_BYTES;
            const int strip_x0 = sx * STRIP_W, y0 = j * TH;
            int* Lw = (int*)sb;
            if (USE_TMA) {
                mbar_wait_bounded(&stage_bar[st], (unsigned)(i / NS) & 1u, P.counters);
            } else {
                __syncwarp();
                for (int k = lane; k < TH * LAB_PITCH; k += 32) {
                    const int r = k / LAB_PITCH, cidx = k - r * LAB_PITCH;
                    const int gy = y0 + r, gx = strip_x0 + cidx;
                    Lw[k] = (gy < P.rows_avail && gx < P.W) ? P.labels[(int64_t)gy * P.ld + gx] : 0;
                }
                if constexpr (C > 0) {
                    unsigned char* Ib = sb + CF::LAB_BOX;
                    for (int k = lane; k < TH * CF::IMG_ROW_WORDS * 4; k += 32) {
                        const int r = k / (CF::IMG_ROW_WORDS * 4), bidx = k - r * (CF::IMG_ROW_WORDS * 4);
                        const int gy = y0 + r;
                        const int64_t gb = (int64_t)strip_x0 * C + bidx;
                        Ib[k] = (gy < P.rows_own && gb < (int64_t)P.W * C) ? P.image[(int64_t)gy * P.image_pitch + gb] : 0;
                    }
                }
                __syncwarp();
            }

            const int x0 = strip_x0 + 4 * lane;
            const bool in_img = x0 < P.W;
            if (!contiguous) {
                up = make_int4(0, 0, 0, 0);
                if (y0 > 0) {
                    const int32_t* row = P.labels + (int64_t)(y0 - 1) * P.ld;
                    if (x0 < P.W) up.x = row[x0];
                    if (x0 + 1 < P.W) up.y = row[x0 + 1];
                    if (x0 + 2 < P.W) up.z = row[x0 + 2];
                    if (x0 + 3 < P.W) up.w = row[x0 + 3];
                }
            }

            // ---- the block row -------------------------------------------------------------------------------------
            const int* Lp = Lw + 4 * lane;
            const int4 a0 = *(const int4*)Lp;
            const int4 a1 = *(const int4*)(Lp + LAB_PITCH);
            const int4 a2 = *(const int4*)(Lp + 2 * LAB_PITCH);
            const int4 a3 = *(const int4*)(Lp + 3 * LAB_PITCH);
            int r0 = __shfl_down_sync(0xffffffffu, a0.x, 1);
            int r1 = __shfl_down_sync(0xffffffffu, a1.x, 1);
            int r2 = __shfl_down_sync(0xffffffffu, a2.x, 1);
            int r3 = __shfl_down_sync(0xffffffffu, a3.x, 1);
            if (lane == 31) {
                r0 = Lp[4];
                r1 = Lp[LAB_PITCH + 4];
                r2 = Lp[2 * LAB_PITCH + 4];
                r3 = Lp[3 * LAB_PITCH + 4];
            }
            const unsigned d = or3(or3(or3(eq3(up.x, up.y, up.z), eq3(up.z, up.w, a0.x), eq3(a0.x, a0.y, a0.z)),
                                       or3(eq3(a0.z, a0.w, r0), eq3(r0, a1.x, a1.y), eq3(a1.y, a1.z, a1.w)),
                                       or3(eq3(a1.w, r1, a2.x), eq3(a2.x, a2.y, a2.z), eq3(a2.z, a2.w, r2))),
                                   or3(eq3(r2, a3.x, a3.y), eq3(a3.y, a3.z, a3.w), eq3(a3.w, r3, up.x)), 0u);
            const bool forced = (x0 == 0) || (x0 + 4 >= P.W) || (y0 == 0) || (y0 + 4 >= P.rows_own);
            const bool is_item = in_img && (forced || d != 0u);
            const bool is_fast = in_img && !is_item;
            const int ref = a0.x;
            if (is_fast && ref != cur) {
                lane_flush();
                cur = ref;
            }
            __syncwarp();
            const unsigned* Ip = (const unsigned*)(sb + CF::LAB_BOX) + CW * lane;
            if (is_fast) {
                area += 16u;
                if constexpr (C > 0) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        unsigned px[CW];
                        if constexpr (C == 4) {
                            const uint4 v = *(const uint4*)(Ip + r * CF::IMG_ROW_WORDS);
                            px[0] = v.x; px[1] = v.y; px[2] = v.z; px[3] = v.w;
                        } else {
#pragma unroll
                            for (int c = 0; c < C; ++c) px[c] = Ip[r * CF::IMG_ROW_WORDS + c];
                        }
                        ragcore::row_stats<C>(px, s, q);
                    }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, is_item);
            if (bal) {
                if (is_item) {
                    const unsigned idx = produced + __popc(bal & lanemask_lt());
                    const unsigned bq = idx >> 5;                       // batch number
                    const unsigned g = bq % NB, gen = bq / NB;
                    wait_consumed(&consumed[g], gen, P.counters);      // the batch's previous generations are consumed
                    uint4* it = ring + (idx % SLOTS);
                    it[0] = make_uint4((unsigned)up.x, (unsigned)up.y, (unsigned)up.z, (unsigned)up.w);
                    it[1 * SLOTS] = make_uint4((unsigned)a0.x, (unsigned)a0.y, (unsigned)a0.z, (unsigned)a0.w);
                    it[2 * SLOTS] = make_uint4((unsigned)a1.x, (unsigned)a1.y, (unsigned)a1.z, (unsigned)a1.w);
                    it[3 * SLOTS] = make_uint4((unsigned)a2.x, (unsigned)a2.y, (unsigned)a2.z, (unsigned)a2.w);
                    it[4 * SLOTS] = make_uint4((unsigned)a3.x, (unsigned)a3.y, (unsigned)a3.z, (unsigned)a3.w);
                    it[5 * SLOTS] = make_uint4((unsigned)r0, (unsigned)r1, (unsigned)r2, (unsigned)r3);
                    if constexpr (C > 0) {
                        unsigned Wf[4 * CW];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
#pragma unroll
                            for (int c = 0; c < C; ++c) Wf[r * CW + c] = Ip[r * CF::IMG_ROW_WORDS + c];
                        }
#pragma unroll
                        for (int vi = 0; vi < C; ++vi)
                            it[(6 + vi) * SLOTS] = make_uint4(Wf[4 * vi], Wf[4 * vi + 1], Wf[4 * vi + 2], Wf[4 * vi + 3]);
                    }
                    it[(6 + C) * SLOTS] = make_uint4((unsigned)x0, (unsigned)y0, 0u, 0u);
                    mbar_arrive_release(&full_b[g], 1u);
                }
                produced += __popc(bal);
            }
            up = a3;
            contiguous = (j + 1 < P.tiles_y);
            if (j + 1 < P.tiles_y) ++j;
            else { j = 0; ++sx; }

            __syncwarp();
            if (USE_TMA && i + NS < my_units) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue();
            }
            ++units_since_flush;
            const bool forced_flush = units_since_flush >= CF::FLUSH_UNITS || (i + 1 == my_units);
            if (forced_flush) {
                lane_flush();
                units_since_flush = 0;
            }
            __syncwarp();
            if (forced_flush || T.used[0] > RS / 2) drain_tables_impl<C>(tab, &P, lane);
        }

        // ---- end of this warp's input: every item has been written ---------------------------------------------------
        __syncwarp();
        if (lane == 0) {
            ctrl->final_total = produced;
            __threadfence_block();
            const unsigned rem = produced & 31u;
            if (rem) mbar_arrive_release(&full_b[(produced >> 5) % NB], 32u - rem);
            __threadfence_block();
            atomicExch(&ctrl->finished, 1u);
        }
    } else {
        // =============================================================================================================
        // ITEM role
        // =============================================================================================================
        unsigned* tab = (unsigned*)(pbase + NS * CF::STAGE_BYTES + CF::TABLE_BYTES);
        const Tab<C> T = Tab<C>::from(tab);
        T.rkey[lane] = EMPTY_LABEL;
        T.rarea[lane] = 0;
        T.rborder[lane] = 0;
#pragma unroll
        for (int cc = 0; cc < C; ++cc) {
            T.rsum[cc * RS + lane] = 0;
            T.rsq[cc * RS + lane] = 0;
        }
        T.ekey[lane] = EMPTY_KEY;
        T.ekey[lane + 32] = EMPTY_KEY;
        T.ecnt[lane] = 0;
        T.ecnt[lane + 32] = 0;
        if (lane == 0) T.used[0] = T.used[1] = 0;
        __syncwarp();
        const ragcore::Geo g{P.W, P.rows_own, P.rows_avail, P.top_border, P.bottom_border};
        const InlineSink<C> sink{T, P};
        int batches_since_drain = 0;

        for (unsigned bq = 0;; ++bq) {
            const unsigned gb = bq % NB, parity = (bq / NB) & 1u;
            bool live = true;
            while (!mbar_try_wait(&full_b[gb], parity)) {
                if (ld_volatile_u32(&ctrl->finished) && (bq << 5) >= ld_volatile_u32(&ctrl->final_total)) {
                    live = false;
                    break;
                }
            }
            if (!live) break;
            const unsigned total = ld_volatile_u32(&ctrl->final_total);
            const unsigned idx = (bq << 5) + (unsigned)lane;
            if (idx < total) {
                const int jslot = (int)(idx % SLOTS);
                const uint4* it = ring + jslot;
                int lab[ragcore::WIN];
                unsigned img[4 * CW];
                uint4 v = it[0];
                lab[20] = (int)v.x; lab[21] = (int)v.y; lab[22] = (int)v.z; lab[23] = (int)v.w;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    v = it[(1 + r) * SLOTS];
                    lab[5 * r] = (int)v.x; lab[5 * r + 1] = (int)v.y; lab[5 * r + 2] = (int)v.z; lab[5 * r + 3] = (int)v.w;
                }
                v = it[5 * SLOTS];
                lab[4] = (int)v.x; lab[9] = (int)v.y; lab[14] = (int)v.z; lab[19] = (int)v.w;
#pragma unroll
                for (int vi = 0; vi < C; ++vi) {
                    v = it[(6 + vi) * SLOTS];
                    img[4 * vi] = v.x; img[4 * vi + 1] = v.y; img[4 * vi + 2] = v.z; img[4 * vi + 3] = v.w;
                }
                v = it[(6 + C) * SLOTS];
                const RingPick<SLOTS> pick{(const int*)ring, jslot};
                ragcore::process_item<C>(lab, img, (int)v.x, (int)v.y, g, sink, pick);
            }
            __syncwarp();
            if (lane == 0) st_release_u32(&consumed[gb], bq / NB + 1u);  // the batch's slots may be overwritten
            ++batches_since_drain;
            // a table slot holds at most 32 items x 16 pixels per batch: drain long before 255^2 * pixels reaches 2^32
            const bool forced = batches_since_drain >= 64;
            if (forced || T.used[0] > RS / 2 || T.used[1] > ES / 2) {
                drain_tables_impl<C>(tab, &P, lane);
                batches_since_drain = 0;
            }
        }
        drain_tables_impl<C>(tab, &P, lane);
    }
}

template <typename CF>
static int launch(const Params& Pin, bool allow_tma, cudaStream_t s) {
    Params P = Pin;
    P.tiles_x = (int)ceil_div(P.W, STRIP_W);
    P.tiles_y = (int)ceil_div(P.rows_avail, CF::TH);
    const long long total = (long long)P.tiles_x * P.tiles_y;
    if (total == 0) return DM_OK;
    const long long max_warps = (long long)num_sms() * CF::NSW;
    const long long per = ceil_div(total, max_warps);
    if (per > 0x7fffffff) return DM_ERR_BAD_ARG;
    P.tiles_per_cta = (int)per;
    const int grid = (int)ceil_div(ceil_div(total, per), CF::NSW);

    CUtensorMap mapL, mapI;
    memset(&mapL, 0, sizeof(mapL));
    memset(&mapI, 0, sizeof(mapI));
    bool tma = allow_tma && ((uintptr_t)P.labels % 16 == 0) && ((P.ld * 4) % 16 == 0) && P.ld >= P.W;
    if (CF::C > 0)
        tma = tma && ((uintptr_t)P.image % 16 == 0) && (P.image_pitch % 16 == 0) && (((int64_t)P.W * CF::C) % 4 == 0);
    if (tma)
        tma = make_map_2d(&mapL, P.labels, (uint64_t)P.W, (uint64_t)P.rows_avail, (uint64_t)P.ld * 4, LAB_PITCH, CF::TH);
    if (tma && CF::C > 0)
        tma = make_map_2d(&mapI, P.image, (uint64_t)P.W * CF::C / 4, (uint64_t)P.rows_own, (uint64_t)P.image_pitch,
                          CF::IMG_ROW_WORDS, CF::TH);
    set_last_path(tma ? 1 : 0);
    if (tma) {
        auto k = rag_split_kernel<CF, true>;
        DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        DM_COUNT_LAUNCH(); k<<<grid, CF::THREADS, CF::SMEM_BYTES, s>>>(mapL, mapI, P);
    } else {
        auto k = rag_split_kernel<CF, false>;
        DM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        DM_COUNT_LAUNCH(); k<<<grid, CF::THREADS, CF::SMEM_BYTES, s>>>(mapL, mapI, P);
    }
    DM_LAUNCH_CHECK();
    return DM_OK;
}

}  // namespace split

// <bands, warp pairs per CTA, batches per ring>
int run_split(const Params& P, int C, bool allow_tma, cudaStream_t s) {
    using namespace split;
    if (C != 4) return run_blocks(P, C, allow_tma, s);
    const char* e = getenv("DM_RAG_CFG");
    const int v = e ? atoi(e) : 0;
    switch (v) {
        case 1: return launch<Cfg<4, 8, 2>>(P, allow_tma, s);
        case 2: return launch<Cfg<4, 7, 3>>(P, allow_tma, s);
        case 3: return launch<Cfg<4, 6, 4>>(P, allow_tma, s);
        default: return launch<Cfg<4, 9, 2>>(P, allow_tma, s);
    }
}

}  // namespace rag
}  // namespace dm
