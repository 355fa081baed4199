// readback.cuh — part of griddy_cuda.cu (one translation unit).  Read-back: ghosts, population count, state formatting.
#pragma once

namespace griddy {

// ------------------------------------------------------------------------------------------------
// Read-back: format the world on the device (by actor id), then one copy per output array.

// Ghost rule for downloads: flag every leader that shares its follower's key.
__global__ void __launch_bounds__(256) k_flag_ghosts(Params p, int par, uint8_t* ghost) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= p.scal[SC_NSLOTS]) return;
  const uint32_t st = p.sa[a].st;
  if ((st & (ST_ALIVE | ST_ONROAD)) != (ST_ALIVE | ST_ONROAD)) return;
  const double* pos = p.pos[par];
  const double cur = pos[a];
  for (int32_t x = p.sa[a].ahead; x >= 0 && pos[x] == cur; x = p.sa[x].ahead) ghost[x] = 1;
}

// Living actors of the world: alive and not a ghost.
__global__ void __launch_bounds__(256) k_count_alive(Params p, const uint8_t* ghost, unsigned long long* out) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  const bool live = a < p.scal[SC_NSLOTS] && (p.sa[a].st & ST_ALIVE) && !ghost[a];
  const uint32_t m = __ballot_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

struct PackOut {
  uint8_t *alive, *loc_kind, *side, *route_status;
  int32_t *container, *agenda_cur, *sleep_left, *route_head, *gid;
  double* pos;
  int by_slot;   // index the outputs by slot (multi-GPU read-back) instead of by actor id
};

__global__ void __launch_bounds__(256) k_pack_state(Params p, int par, const uint8_t* ghost, PackOut o) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  if (a >= p.scal[SC_NSLOTS]) return;
  const uint32_t st = p.sa[a].st;
  const int32_t id = o.by_slot ? a : p.cold[a].gid;
  const int head = (st >> ST_HEAD_SHIFT) & 3, rs = (st >> ST_RS_SHIFT) & 3;
  const int32_t aux = p.aux[a];
  if (o.gid) o.gid[id] = p.cold[a].gid;
  if (o.alive) o.alive[id] = ((st & ST_ALIVE) && !ghost[a]) ? 1 : 0;
  if (o.loc_kind) o.loc_kind[id] = (st & ST_ONROAD) ? 1 : 0;
  if (o.container) o.container[id] = p.cold[a].container;
  if (o.side) o.side[id] = (!(st & ST_ONROAD) && (st & ST_SIDE)) ? 1 : 0;
  if (o.pos) o.pos[id] = p.pos[par][a];
  if (o.agenda_cur) o.agenda_cur[id] = p.cold[a].agenda_cur - p.coldf[a].agenda_beg;
  if (o.sleep_left) o.sleep_left[id] = head == HEAD_SLEEP ? aux : 0;
  if (o.route_status) o.route_status[id] = (uint8_t)rs;
  if (o.route_head) o.route_head[id] = rs == GRIDDY_ROUTE_SOME ? p.cold[a].hop : -2;
}

}  // namespace griddy
