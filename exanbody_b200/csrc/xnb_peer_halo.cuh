// xnb_peer_halo.cuh -- the halo exchange over NVLink peer memory: pack + transfer + signal in ONE kernel, no NCCL on the step path.
//
// reference: mpi/include/exanb/mpi/update_ghosts.h / update_ghosts_comm_manager.h:267-281,390,436 (pack -> MPI_Isend / MPI_Irecv per
//            partner -> wait -> unpack), update_ghost_functors.h:41-67,172-217 (pack), :369-452 (unpack), and
//            mpi/particle_displ_over.cu:174 (MPI_Allreduce(SUM) of the displacement count every step).
//
// Every rank owns a MAILBOX in its own HBM (cudaMalloc, exported once with cudaIpcGetMemHandle, mapped by its partners with
// cudaIpcOpenMemHandle: NVLink 5 / NVSwitch peer stores):
//   bytes [0, 4096)          PeerHdr: over[2][64] (displacement counts, one slot per sender and epoch parity), flag[64] (epoch of the
//                            last complete slab of each sender), done (block counter of my own push kernel -- local use only)
//   bytes [4096, ...)        two halves of cap_words 8-byte words: half (epoch & 1) receives the slabs of exchange number `epoch`,
//                            laid out exactly like the NCCL path's receive staging (one slab per partner, field-major)
// k_ghost_push   = k_ghost_pack whose stores go straight into the PARTNER's mailbox at the offset the partner's unpack expects; the
//                  last block to finish publishes flag[me] = epoch on every partner (release, system scope).
// k_ghost_pull   = k_ghost_unpack that first waits (acquire, system scope) for the flags of the partners its elements come from.
// k_peer_allsum  = the all-reduce of the displacement counter: every rank stores (epoch, count) into slot [me] of every mailbox and
//                  sums the nranks slots of its own.
// Why two halves and no back-signal: exchange e + 2 reuses the half of exchange e.  A sender issues push(e + 2) after its own
// pull(e + 1) (stream order), which waited for the receiver's push(e + 1), which the receiver issued after its pull(e): the half is
// free.  The same argument covers over[][].  Flags only grow, so a fresh (zeroed) mailbox after a capacity change needs no reset.
// A wait that sees nothing for PEER_TIMEOUT_NS sets DERR_PEER_TIMEOUT and gives up (a dead partner must not hang the GPU).
#pragma once
#include "xnb_kernels.cuh"

namespace xnb {

constexpr int PEER_MAX_RANKS = 64;
constexpr size_t PEER_HDR_BYTES = 4096;
constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;      // default; XNB_PEER_TIMEOUT_MS overrides (tests)
struct PeerHdr
{
  unsigned long long over[2][PEER_MAX_RANKS];
  unsigned long long flag[PEER_MAX_RANKS];
  unsigned int done; unsigned int pad;
};
static_assert(sizeof(PeerHdr) <= PEER_HDR_BYTES, "mailbox header");
// what a rank knows about partner p's mailbox: where it is mapped here, its half size, and where MY slab starts in p's receive layout
struct PeerSlot { unsigned long long base; unsigned long long cap_words; uint32_t off; uint32_t pad; };

XNB_DEVINL unsigned long long ld_acquire_sys(const unsigned long long* p)
{
  unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
XNB_DEVINL void st_release_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory"); }
XNB_DEVINL unsigned long long global_timer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// spin until *p >= want (MATCH: until the high 32 bits equal want); false on timeout
template <bool MATCH>
XNB_DEVINL bool peer_wait(const unsigned long long* p, unsigned long long want, unsigned long long* got, unsigned long long timeout_ns)
{
  unsigned long long t0 = 0; unsigned int spins = 0;
  for (;;)
  {
    const unsigned long long v = ld_acquire_sys(p);
    if (MATCH ? ((v >> 32) == want) : (v >= want)) { if (got) *got = v; return true; }
    if (++spins > 64u)
    {
      __nanosleep(200);
      if ((spins & 255u) == 0u) { const unsigned long long t = global_timer_ns(); if (t0 == 0) t0 = t; else if (t - t0 > timeout_ns) return false; }
    }
  }
}

template <bool ALL_FIELDS>
__global__ void __launch_bounds__(256)
k_ghost_push(GridP g, int n_send, const uint32_t* __restrict__ send_src, const uint16_t* __restrict__ send_flags,
             ParticlesP p, int self_first, int self_end, uint32_t self_dst,
             const uint32_t* __restrict__ send_base, int nranks, int me, const PeerSlot* __restrict__ peers, unsigned long long epoch, PeerHdr* my_hdr)
{
  constexpr size_t NW = ALL_FIELDS ? GHOST_WORDS_ALL : GHOST_WORDS_R;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_send; q += gridDim.x * blockDim.x)
  {
    const uint32_t s = send_src[q], fl = send_flags[q];
    const double x = coord_shift(p.rx[s], g.dmin[0], g.dmax[0], fl >> 0);
    const double y = coord_shift(p.ry[s], g.dmin[1], g.dmax[1], fl >> 3);
    const double z = coord_shift(p.rz[s], g.dmin[2], g.dmax[2], fl >> 6);
    if (q >= self_first && q < self_end)
    {
      const uint32_t d = self_dst + (uint32_t)(q - self_first);
      p.rx[d] = x; p.ry[d] = y; p.rz[d] = z;
      if (ALL_FIELDS)
      {
        p.vx[d] = p.vx[s]; p.vy[d] = p.vy[s]; p.vz[d] = p.vz[s];
        p.fx[d] = p.fx[s]; p.fy[d] = p.fy[s]; p.fz[d] = p.fz[s];
        p.id[d] = p.id[s]; p.type[d] = p.type[s];
      }
    }
    else
    {
      const int pr = ghost_partner_of(send_base, nranks, (uint32_t)q);
      const size_t s0 = send_base[pr], n = send_base[pr + 1] - s0;
      const PeerSlot ps = peers[pr];
      double* slab = reinterpret_cast<double*>(ps.base + PEER_HDR_BYTES) + (epoch & 1ull) * ps.cap_words + NW * (size_t)ps.off + ((size_t)q - s0);
      slab[0] = x; slab[n] = y; slab[2 * n] = z;
      if (ALL_FIELDS)
      {
        slab[3 * n] = p.vx[s]; slab[4 * n] = p.vy[s]; slab[5 * n] = p.vz[s];
        slab[6 * n] = p.fx[s]; slab[7 * n] = p.fy[s]; slab[8 * n] = p.fz[s];
        reinterpret_cast<unsigned long long*>(slab)[9 * n] = p.id[s];
        reinterpret_cast<unsigned long long*>(slab)[10 * n] = (unsigned long long)p.type[s];
      }
    }
  }
  // publish: every thread's peer stores are ordered before its block's ticket; the block that takes the last ticket raises the flags
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(&my_hdr->done, 1u) == gridDim.x - 1u);
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x == 0) my_hdr->done = 0u;
  for (int pr = threadIdx.x; pr < nranks; pr += blockDim.x)
    if (pr != me && send_base[pr + 1] > send_base[pr])
      st_release_sys(&reinterpret_cast<PeerHdr*>(peers[pr].base)->flag[me], epoch);
}

template <bool ALL_FIELDS>
__global__ void __launch_bounds__(256)
k_ghost_pull(int n_ghost, uint32_t n_inner, ParticlesP p, const uint32_t* __restrict__ recv_base, int nranks, int self_rank,
             const PeerHdr* my_hdr, const double* half, unsigned long long epoch, uint32_t* __restrict__ err, unsigned long long timeout_ns)
{
  constexpr size_t NW = ALL_FIELDS ? GHOST_WORDS_ALL : GHOST_WORDS_R;
  __shared__ int s_lo, s_hi; __shared__ bool s_ok;
  int seen_lo = 0, seen_hi = -1;                       // partners this block has already waited for
  if (threadIdx.x == 0) s_ok = true;
  for (int first = blockIdx.x * blockDim.x; first < n_ghost; first += gridDim.x * blockDim.x)
  {
    const int lastq = min(first + (int)blockDim.x, n_ghost) - 1;
    __syncthreads();
    if (threadIdx.x == 0)
    {
      const int lo = ghost_partner_of(recv_base, nranks, (uint32_t)first), hi = ghost_partner_of(recv_base, nranks, (uint32_t)lastq);
      for (int pr = lo; pr <= hi && s_ok; pr++)
      {
        if (pr == self_rank || recv_base[pr + 1] == recv_base[pr] || (pr >= seen_lo && pr <= seen_hi)) continue;
        if (!peer_wait<false>(&my_hdr->flag[pr], epoch, nullptr, timeout_ns)) { s_ok = false; atomicOr(err, DERR_PEER_TIMEOUT); }
      }
      if (seen_hi < seen_lo) seen_lo = lo;
      seen_hi = hi;
      s_lo = lo; s_hi = hi;
    }
    __syncthreads();
    if (!s_ok) return;
    const int i = first + (int)threadIdx.x;
    if (i > lastq) continue;
    const int pr = (s_lo == s_hi) ? s_lo : ghost_partner_of(recv_base, nranks, (uint32_t)i);
    if (pr == self_rank) continue;                     // my own periodic images were written in place by the push kernel
    const size_t r0 = recv_base[pr], n = recv_base[pr + 1] - r0;
    const double* slab = half + NW * r0 + ((size_t)i - r0);
    const uint32_t d = n_inner + (uint32_t)i;
    // written by another GPU: read through L2 (.cg), never a stale L1 line
    p.rx[d] = __ldcg(slab); p.ry[d] = __ldcg(slab + n); p.rz[d] = __ldcg(slab + 2 * n);
    if (ALL_FIELDS)
    {
      p.vx[d] = __ldcg(slab + 3 * n); p.vy[d] = __ldcg(slab + 4 * n); p.vz[d] = __ldcg(slab + 5 * n);
      p.fx[d] = __ldcg(slab + 6 * n); p.fy[d] = __ldcg(slab + 7 * n); p.fz[d] = __ldcg(slab + 8 * n);
      p.id[d] = __ldcg(reinterpret_cast<const unsigned long long*>(slab) + 9 * n);
      p.type[d] = (uint8_t)__ldcg(reinterpret_cast<const unsigned long long*>(slab) + 10 * n);
    }
  }
}

// sum over ranks of value[0] (in place).  One block, one thread per rank (nranks <= PEER_MAX_RANKS).
__global__ void __launch_bounds__(PEER_MAX_RANKS)
k_peer_allsum(int nranks, int me, const PeerSlot* __restrict__ peers, const PeerHdr* my_hdr, unsigned long long epoch, unsigned long long* value,
              uint32_t* __restrict__ err, unsigned long long timeout_ns)
{
  __shared__ unsigned long long s_part[PEER_MAX_RANKS];
  const int t = threadIdx.x;
  const unsigned long long mine = min(*value, 0xffffffffull);
  const unsigned long long e32 = epoch & 0xffffffffull;
  unsigned long long got = 0;
  if (t < nranks)
  {
    st_release_sys(&reinterpret_cast<PeerHdr*>(peers[t].base)->over[epoch & 1ull][me], (e32 << 32) | mine);
    if (!peer_wait<true>(&my_hdr->over[epoch & 1ull][t], e32, &got, timeout_ns)) { atomicOr(err, DERR_PEER_TIMEOUT); got = 1; }     // a timeout forces the rebuild path
  }
  s_part[t] = (t < nranks) ? (got & 0xffffffffull) : 0ull;
  __syncthreads();
  if (t == 0)
  {
    unsigned long long s = 0;
    for (int q = 0; q < nranks; q++) s += s_part[q];
    *value = s;
  }
}

} // namespace xnb
