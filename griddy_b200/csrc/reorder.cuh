// reorder.cuh — part of griddy_cuda.cu (one translation unit).  Periodic spatial reorder of the slots (with radix_sort.cuh).
#pragma once

namespace griddy {

// ------------------------------------------------------------------------------------------------
// Reorder: sort the live slots by (locality key of the container, coarse pos-param), permute every
// per-slot array, remap the slot-valued pointers (ahead, behind, lane tails) and drop the dead.

struct Permute {
  const Hot* src_hot;
  Hot* dst_hot;
  const double* src_pos;
  double* dst_pos;
  const Cold* src_cold;
  Cold* dst_cold;
  const ColdF* src_coldf;
  ColdF* dst_coldf;
  AgendaItem* agenda_dst;   // partitioned worlds: the pool the remaining agenda items are compacted into (else null)
  int32_t* n_items_new;     //   and its fill count
};

// grid = n_pad / 256 exactly: every lane of every warp reaches the ballot
__global__ void __launch_bounds__(256) k_make_keys(Params p, int par, int key_bits, uint64_t* __restrict__ keys,
                                                    uint32_t* __restrict__ vals, uint8_t* __restrict__ tailflag) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  uint64_t key = ~0ull;
  uint32_t val = 0xffffffffu;
  bool live = false;
  if (s < p.scal[SC_NSLOTS]) {
    const uint32_t st = p.hot[s].st;
    if (st & ST_ALIVE) {
      live = true;
      const int32_t c = p.cold[s].container;
      const double x = p.pos[par][s];
      // 16 bits of pos-param over [-0.5, 1.5): enough to keep a lane's actors in list order
      const double q = fmin(fmax((x + 0.5) * 32768.0, 0.0), 65535.0);
      // off-road actors after the on-road ones: warps are then all-travelling or all-parked
      key = ((uint64_t)((st & ST_ONROAD) ? 0u : 1u) << (key_bits - 1)) | ((uint64_t)p.loc_key[c] << 16) | (uint64_t)(uint32_t)q;
      val = (uint32_t)s;
      tailflag[s] = ((st & ST_ONROAD) && p.lane[c].tail == s) ? 1 : 0;
    }
  }
  keys[s] = key;
  vals[s] = val;
  const uint32_t m = __ballot_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(&p.scal[SC_NALIVE], __popc(m));
}

__global__ void __launch_bounds__(256) k_inverse(const uint32_t* __restrict__ vals, int n_pad, int32_t* __restrict__ newslot) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= n_pad) return;
  const uint32_t old = vals[s];
  if (old != 0xffffffffu) newslot[old] = s;
}

__global__ void __launch_bounds__(256) k_permute(Params p, Permute q, const uint32_t* __restrict__ vals,
                                                  const int32_t* __restrict__ newslot,
                                                  const uint8_t* __restrict__ tailflag) {
  const int s = blockIdx.x * 256 + threadIdx.x;
  if (s >= p.scal[SC_NALIVE]) return;
  const uint32_t old = vals[s];
  // ahead / behind are meaningful on the road only; an off-road actor's values are stale
  Hot h = q.src_hot[old];
  const bool on_road = h.st & ST_ONROAD;
  if (on_road && h.ahead >= 0) {
    h.ahead = newslot[h.ahead];
    if (h.ahead < 0) dev_error(p, DEV_ERR_INTERNAL);   // a live actor's leader must be live
  } else {
    h.ahead = -1;
  }
  if (on_road && h.behind >= 0) {
    h.behind = newslot[h.behind];
    if (h.behind < 0) dev_error(p, DEV_ERR_INTERNAL);   // a live actor's follower must be live
  } else {
    h.behind = -1;
  }
  q.dst_hot[s] = h;
  q.dst_pos[s] = q.src_pos[old];
  Cold cold = q.src_cold[old];
  ColdF coldf = q.src_coldf[old];
  if (q.agenda_dst) {   // only the items not yet popped are kept; agenda_beg stays "head minus items popped"
    const int32_t rem = coldf.agenda_end - cold.agenda_cur, popped = cold.agenda_cur - coldf.agenda_beg;
    const int32_t base = rem > 0 ? atomicAdd(q.n_items_new, rem) : 0;
    for (int32_t k = 0; k < rem; ++k) q.agenda_dst[base + k] = p.agenda[cold.agenda_cur + k];
    cold.agenda_cur = base;
    coldf.agenda_beg = base - popped;
    coldf.agenda_end = base + rem;
  }
  q.dst_cold[s] = cold;
  q.dst_coldf[s] = coldf;
  if (tailflag[old]) p.lane[cold.container].tail = s;
}

}  // namespace griddy
