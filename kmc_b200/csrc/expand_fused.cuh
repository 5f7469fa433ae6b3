// kmc_b200 — super-k-mer expansion in ONE pass over the bin (the default path of the bin pipeline; expand.cuh keeps the
// index-based kernels for oversized bins and for packs of more than 64 KiB).
//
// Replaces CKmerBinSorter::ExpandKmersBoth / ExpandKmersAll (kmc_core/kb_sorter.h:251-362; for k % 32 != 0 also ExpandKxmers*
// :371-724: we always expand to plain k-mers) and removes what the index-based path needed on top: the per-super-k-mer index
// (8 bytes per super-k-mer written and read back), the pack scan kernel, and the per-tile metadata chain of expand_kernel.
//
// One CTA per expander pack (<= 64 KiB = one collector flush, kb_collector.cpp:93-106), one THREAD per 256-byte segment of it:
//   1. the pack is staged in shared memory as big-endian 32-bit words, one pad word per 256 bytes (so that the 32 lanes of a warp,
//      which work 256 bytes apart, hit 32 different banks);
//   2. the record chain ("the length byte says where the next record starts") is walked by all segments in parallel from speculative
//      starts, repaired and verified exactly as in walk_packs_parallel_kernel (expand.cuh);
//   3. the pack's first output index comes from a decoupled look-back over the packs' k-mer counts (no scan kernel, no second launch);
//   4. every thread then EXPANDS ITS OWN SEGMENT sequentially: first k-mer of a record by funnel shifts, every further one by rolling
//      one symbol in (forward strand: shift left, reverse complement: shift right - the reference's own loop, kb_sorter.h:343-358),
//      canonical = min(kmer, rev).  The k-mers pass through a small per-lane queue in shared memory and leave in runs of up to 8
//      consecutive records per lane, so every 32-byte sector is written once;
//   5. the level-1 digit counts of the MSD sort (msd_sort.cuh) are accumulated per OUTPUT TILE (aligned tiles of msd_tile() records)
//      in shared memory and added to the global cells at the end (32-bit atomics on pairs of u16 cells: only the two boundary tiles
//      of a pack are shared with its neighbours).
// Traffic: the bin bytes once in, the records once out (+ the cells).
#pragma once
#include "common.cuh"
#include "expand.cuh"

namespace kmcb {

constexpr int kFxThreads = kWalkSegs;            // 256: one thread per 256-byte segment
constexpr int kFxTiles = 24;                     // output tiles whose digit counts are held in shared memory at a time (a 64 KiB pack of
                                                 // ~12-k-mer super-k-mers spans <= 18 tiles of 4096 one-word records; longer packs take another window)
template <int WORDS> struct FxCfg {
	static constexpr int kQ = WORDS == 1 ? 8 : WORDS == 2 ? 4 : 2;          // queued records per lane between two flushes
	static constexpr int kLaneStrideWords = kQ * 2 * WORDS + 2;             // 32-bit words between two lanes' queues (+2: conflict-free 8-byte stores)
};
constexpr int kFxPackWords = (kWalkChunk + 64) / 4;                         // staged words (the 16-byte grid shifts the pack by up to 15 bytes; the funnel shifts look <= 2*WORDS+1 words ahead)
constexpr int kFxPaddedWords = kFxPackWords + kFxPackWords / 64 + 2;

template <int WORDS>
struct FxSmem {
	static constexpr int oPack = 0;
	static constexpr int oHtop = ((kFxPaddedWords * 4) + 15) & ~15;
	static constexpr int oQueue = oHtop + kFxTiles * 256 * 4;
	static constexpr int oEntry = oQueue + (kFxThreads / 32) * 32 * FxCfg<WORDS>::kLaneStrideWords * 4;
	static constexpr int oExit = oEntry + kFxThreads * 4;
	static constexpr int kBytes = oExit + kFxThreads * 4;
};

struct FusedArgs {
	const uint8_t* bin;
	uint64_t size;
	const uint64_t* pack_start;  // [n_packs + 1] byte offsets (device)
	uint32_t n_packs;
	uint32_t k;
	uint32_t both_strands;
	uint64_t n_rec;              // k-mers the bin must hold (CBinDesc::n_rec): the record buffer is sized from it, nothing is written beyond
	void* recs;
	uint32_t* cells1;            // level-1 cells u16 [256][n_tiles] seen as 32-bit words, zero-initialised
	uint32_t n_tiles;            // ceil(n_rec / tile)
	uint32_t tile_shift;         // log2(msd_tile<WORDS>())
	uint32_t top_shift;
	uint64_t* desc;              // [n_packs] look-back descriptors (epoch-tagged, common.cuh)
	uint32_t epoch;
	uint32_t* ticket;            // zero-initialised: packs are taken in order
	uint32_t* status;            // [0] error bits (expand.cuh)
	uint32_t* flags;             // [0], [1]: kExpandAbortFlag when the bin is malformed
};

// word `w` of the staged stream (one pad word per 64)
__device__ __forceinline__ uint32_t fx_ldw(const uint32_t* sw, uint32_t w) { return sw[w + (w >> 6)]; }
// byte `i` of the staged stream (big-endian words)
__device__ __forceinline__ uint32_t fx_ldb(const uint32_t* sw, uint32_t i)
{
	const uint32_t w = i >> 2;
	return reinterpret_cast<const uint8_t*>(sw)[4u * (w + (w >> 6)) + (3u - (i & 3u))];
}

template <int WORDS>
__global__ void __launch_bounds__(kFxThreads, 2) expand_fused_kernel(const FusedArgs a)
{
	using S = FxSmem<WORDS>;
	using R = Rec<WORDS>;
	constexpr int Q = FxCfg<WORDS>::kQ, LSW = FxCfg<WORDS>::kLaneStrideWords;
	constexpr uint32_t FULL = 0xffffffffu;
	extern __shared__ __align__(16) uint8_t fx_smem[];
	uint32_t* sw = reinterpret_cast<uint32_t*>(fx_smem + S::oPack);
	uint32_t* htop = reinterpret_cast<uint32_t*>(fx_smem + S::oHtop);
	uint32_t* queue = reinterpret_cast<uint32_t*>(fx_smem + S::oQueue);
	uint32_t* s_entry = reinterpret_cast<uint32_t*>(fx_smem + S::oEntry);
	uint32_t* s_exit = reinterpret_cast<uint32_t*>(fx_smem + S::oExit);
	__shared__ uint32_t s_w[16];
	__shared__ uint32_t s_bad, s_pack, s_fatal;
	__shared__ unsigned long long s_kbase;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

	if (tid == 0) { s_pack = atomicAdd(a.ticket, 1u); s_bad = 0; s_fatal = 0; }
	for (uint32_t j = tid; j < (uint32_t)kFxTiles * 256u; j += kFxThreads) htop[j] = 0;
	__syncthreads();
	const uint32_t p = s_pack;
	if (p >= a.n_packs) return;
	const uint64_t pstart = a.pack_start[p];
	const uint64_t plen64 = a.pack_start[p + 1] - pstart;
	const uint32_t len = (uint32_t)min(plen64, (uint64_t)kWalkChunk);           // (the host only takes this path when every pack has <= 64 KiB)

	// ---- 1. the pack into shared memory: 16-byte loads on the absolute 16-byte grid, big-endian words, one pad word per 64
	const uintptr_t g0 = reinterpret_cast<uintptr_t>(a.bin + pstart);
	const uintptr_t g0a = g0 & ~(uintptr_t)15;
	const uint32_t shift = (uint32_t)(g0 - g0a);                                // stream byte i lives at staged byte i + shift
	const uint32_t end = shift + len;                                           // staged end of the pack
	{
		const uint32_t nvec = (end + 15) >> 4;
		for (uint32_t v0 = tid; v0 < nvec; v0 += 8 * kFxThreads) {
			uint4 r[8];
#pragma unroll
			for (int i = 0; i < 8; ++i) { const uint32_t v = v0 + i * kFxThreads; r[i] = v < nvec ? __ldg(reinterpret_cast<const uint4*>(g0a) + v) : make_uint4(0, 0, 0, 0); }
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				const uint32_t v = v0 + i * kFxThreads;
				if (v < nvec) {
					const uint32_t w = 4u * v, o = w + (w >> 6);            // (a 16-byte vector never straddles a 256-byte boundary)
					sw[o] = __byte_perm(r[i].x, 0, 0x0123); sw[o + 1] = __byte_perm(r[i].y, 0, 0x0123);
					sw[o + 2] = __byte_perm(r[i].z, 0, 0x0123); sw[o + 3] = __byte_perm(r[i].w, 0, 0x0123);
				}
			}
		}
		// (words behind the pack: read by the funnel shifts of its last k-mers, their bits are shifted out; keep them defined)
		for (uint32_t w = ((end + 15) >> 4) * 4 + tid; w < ((end + 15) >> 4) * 4 + 2 * WORDS + 2 && w < (uint32_t)kFxPackWords; w += kFxThreads) sw[w + (w >> 6)] = 0;
	}
	__syncthreads();

	// ---- 2. parallel speculative walk (see walk_packs_parallel_kernel): segment t = staged bytes [256 t, 256 t + 256); the last one runs to the end
	const uint32_t k3 = a.k + 3;
	const uint32_t seg_lo = max(tid * (uint32_t)kWalkSegBytes, shift);
	const uint32_t seg_hi = tid == kFxThreads - 1 ? end : min((tid + 1) * (uint32_t)kWalkSegBytes, end);
	const bool has_seg = seg_lo < seg_hi;
	uint32_t pos, entry, nrec = 0, nk = 0;
	{
		pos = seg_lo > shift + (uint32_t)kWalkSpec ? seg_lo - kWalkSpec : shift;
		if (has_seg) { while (pos < seg_lo) pos += 1 + ((fx_ldb(sw, pos) + k3) >> 2); }
		else pos = end;
		entry = min(pos, end);
		pos = entry;
		while (pos < seg_hi) { const uint32_t x = fx_ldb(sw, pos); nk += x + 1; ++nrec; pos += 1 + ((x + k3) >> 2); }
		s_entry[tid] = entry; s_exit[tid] = has_seg ? pos : end;
		for (int it = 0; it < kFxThreads; ++it) {
			__syncthreads();
			const uint32_t want = tid > 0 ? min(s_exit[tid - 1], end) : shift;
			const bool fix = tid > 0 && has_seg && entry != want;
			if (!__syncthreads_or(fix)) break;
			if (fix) {
				entry = want; nrec = 0; nk = 0; pos = entry;
				while (pos < seg_hi) { const uint32_t x = fx_ldb(sw, pos); nk += x + 1; ++nrec; pos += 1 + ((x + k3) >> 2); }
				s_entry[tid] = entry; s_exit[tid] = pos;
			}
		}
		__syncthreads();
		bool bad = false;
		if (tid > 0 && has_seg && s_entry[tid] != min(s_exit[tid - 1], end)) bad = true;
		if (has_seg && seg_hi == end && s_exit[tid] != end) bad = true;          // the last record must end with the pack
		if (plen64 > (uint64_t)kWalkChunk) bad = true;
		if (bad) atomicOr(&s_bad, 1u);
	}
	// exclusive scan of the segments' k-mer counts
	uint32_t ik = nk;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, ik, o); if (lane >= (uint32_t)o) ik += t; }
	if (lane == 31) s_w[warp] = ik;
	__syncthreads();
	uint32_t bk = ik - nk, tk = 0;
#pragma unroll
	for (int w = 0; w < kFxThreads / 32; ++w) { if ((uint32_t)w < warp) bk += s_w[w]; tk += s_w[w]; }
	const bool pack_bad = s_bad != 0;

	// ---- 3. first output index of the pack: decoupled look-back over the packs (every pack publishes, also a malformed one)
	if (tid == 0) {
		const uint64_t excl = lookback_exclusive(a.desc, 1, p, pack_bad ? 0ull : (uint64_t)tk, a.epoch);
		s_kbase = excl;
		bool fatal = pack_bad;
		if (pack_bad) atomicOr(a.status, kErrPackWalk);
		if (!pack_bad && excl + tk > a.n_rec) { atomicOr(a.status, kErrRecCount); fatal = true; }                 // more k-mers than the caller sized the buffers for
		if (!pack_bad && p == a.n_packs - 1 && excl + tk != a.n_rec) { atomicOr(a.status, kErrRecCount); fatal = true; }
		if (fatal) { atomicOr(&a.flags[0], kExpandAbortFlag); atomicOr(&a.flags[1], kExpandAbortFlag); s_fatal = 1; }
	}
	__syncthreads();
	if (pack_bad || (s_fatal && (uint64_t)s_kbase + tk > a.n_rec)) return;      // (a short total is only detected by the last pack: its records are in bounds)
	const uint32_t kbase = (uint32_t)s_kbase;                                   // n_rec < 2^32 (checked on the host)

	// ---- 4. every thread expands its own segment
	R* __restrict__ out = reinterpret_cast<R*>(a.recs);
	const uint32_t k = a.k;
	const uint32_t rs = 64u * WORDS - 2u * k;                                   // right-alignment shift, 0..63
	const uint64_t topmask = (2u * k) % 64u ? ((1ull << ((2u * k) % 64u)) - 1ull) : ~0ull;
	const uint32_t comp_shift = (2u * k - 2u) & 63u;
	const bool canonical = a.both_strands != 0;
	uint32_t* myq = queue + (warp * 32 + lane) * LSW;
	uint32_t idx = kbase + bk;                                                   // output index of my next k-mer
	const uint32_t idx_end = idx + nk;
	uint32_t rem = 0;                                                            // k-mers still to come from the current record
	uint32_t bit = 0, win = 0;                                                   // staged bit position of the next symbol, the word that holds it
	R f, r;
#pragma unroll
	for (int i = 0; i < WORDS; ++i) { f.w[i] = 0; r.w[i] = 0; }
	pos = entry;
	for (uint32_t w0 = kbase >> a.tile_shift;; w0 += kFxTiles) {
		const uint64_t win_end64 = ((uint64_t)(w0 + kFxTiles)) << a.tile_shift;
		const uint32_t win_end = (uint32_t)min(win_end64, (uint64_t)0xffffffffu);
		for (;;) {
			uint32_t cnt = 0;
			const uint32_t base = idx;
#pragma unroll 1
			for (int q = 0; q < Q; ++q) {
				if (idx < idx_end && idx < win_end) {
					if (rem == 0) {
						// a new record: its first k-mer straight from the staged words (as extract_kmer_be32)
						const uint32_t x = fx_ldb(sw, pos);
						rem = x + 1;
						const uint32_t B = 8u * (pos + 1u);
						const uint32_t wi = B >> 5, bo = B & 31u;
						uint32_t w[2 * WORDS + 1];
#pragma unroll
						for (int i = 0; i <= 2 * WORDS; ++i) w[i] = fx_ldw(sw, wi + i);
						uint64_t t[WORDS];
#pragma unroll
						for (int i = 0; i < WORDS; ++i)
							t[i] = ((uint64_t)__funnelshift_l(w[2 * i + 1], w[2 * i], bo) << 32) | (uint64_t)__funnelshift_l(w[2 * i + 2], w[2 * i + 1], bo);
#pragma unroll
						for (int i = 0; i < WORDS; ++i) {
							uint64_t v = t[i] >> rs;
							if (i > 0 && rs) v |= t[i - 1] << (64u - rs);
							f.w[WORDS - 1 - i] = v;
						}
						if (canonical) {
							uint64_t u[WORDS];
#pragma unroll
							for (int i = 0; i < WORDS; ++i) u[i] = ~rev_symbols64(f.w[i]);
#pragma unroll
							for (int i = 0; i < WORDS; ++i) {
								uint64_t v = u[i] >> rs;
								if (i > 0 && rs) v |= u[i - 1] << (64u - rs);
								r.w[WORDS - 1 - i] = v;
							}
						}
						bit = B + 2u * k;
						win = fx_ldw(sw, bit >> 5);
						pos += 1 + ((x + k3) >> 2);
					} else {
						// the next k-mer of the record: one symbol rolled in (kb_sorter.h:343-358)
						if ((bit & 31u) == 0) win = fx_ldw(sw, bit >> 5);
						const uint32_t sym = (win >> (30u - (bit & 31u))) & 3u;
						bit += 2;
#pragma unroll
						for (int i = WORDS - 1; i > 0; --i) f.w[i] = (f.w[i] << 2) | (f.w[i - 1] >> 62);
						f.w[0] = (f.w[0] << 2) | sym;
						f.w[WORDS - 1] &= topmask;
						if (canonical) {
#pragma unroll
							for (int i = 0; i < WORDS - 1; ++i) r.w[i] = (r.w[i] >> 2) | (r.w[i + 1] << 62);
							r.w[WORDS - 1] = (r.w[WORDS - 1] >> 2) | ((uint64_t)(3u - sym) << comp_shift);
						}
					}
					--rem;
					const R o = (!canonical || rec_less<WORDS>(f, r)) ? f : r;           // kmer < rev ? kmer : rev  (kb_sorter.h:340,356)
#pragma unroll
					for (int i = 0; i < WORDS; ++i) *reinterpret_cast<uint64_t*>(myq + (q * WORDS + i) * 2) = o.w[i];
					atomicAdd(&htop[((idx >> a.tile_shift) - w0) * 256u + rec_top_digit<WORDS>(o, a.top_shift)], 1u);
					++idx;
					++cnt;
				}
			}
			// flush: 32 / Q lanes' queues per store instruction, every lane's records are consecutive in the output
			const uint32_t any = __ballot_sync(FULL, cnt != 0);
			if (any == 0) break;
			__syncwarp();
#pragma unroll
			for (int it = 0; it < Q; ++it) {
				const uint32_t L = it * (32 / Q) + lane / Q, e = lane % Q;
				const uint32_t n = __shfl_sync(FULL, cnt, L), b = __shfl_sync(FULL, base, L);
				if (e < n) {
					const uint32_t* src = queue + (warp * 32 + L) * LSW + e * WORDS * 2;
					R o;
#pragma unroll
					for (int i = 0; i < WORDS; ++i) o.w[i] = *reinterpret_cast<const uint64_t*>(src + 2 * i);
					out[b + e] = o;
				}
			}
			__syncwarp();
		}
		// ---- 5. the window's digit counts -> the level-1 cells (pairs of u16 in one 32-bit word; a count never exceeds the tile size)
		__syncthreads();
		for (uint32_t j = tid; j < (uint32_t)kFxTiles * 256u; j += kFxThreads) {
			const uint32_t c = htop[j];
			if (c) {
				htop[j] = 0;
				const uint32_t t = w0 + (j >> 8), d = j & 255u;
				if (t < a.n_tiles) {
					const uint64_t cell = (uint64_t)d * a.n_tiles + t;
					atomicAdd(a.cells1 + (cell >> 1), c << (16u * (uint32_t)(cell & 1ull)));
				}
			}
		}
		const bool more = idx < idx_end;
		if (!__syncthreads_or(more)) break;
	}
}

}  // namespace kmcb
