// kmc_b200 — sorted-run counting and compaction (replaces CKmerBinSorter::CompactKmers,
// kmc_core/kb_sorter.h:1128-1281, and for k % 32 != 0 the (k,x)-mer merge CompactKxmers :937-1122 +
// kxmer_set.h, whose output is the same function of the sorted k-mer multiset).
//
// One pass over the sorted records, one CTA per tile, chained scan (decoupled look-back) for the output
// offsets, so the kernel reads N*W bytes once and writes only the emitted database records:
//   * run tails are found with one neighbour compare per record; ballots give a bitmap of tails, the
//     previous set bit gives the run head -> run length, no scan needed; a run that started in an earlier
//     tile is measured by a short backward probe of the (sorted) input followed, for giant runs, by a
//     lower_bound binary search;
//   * cutoffs and clamping exactly as kb_sorter.h:1174-1191: count < cutoff_min -> n_cutoff_min,
//     else count > cutoff_max -> n_cutoff_max, else emit min(count, counter_max);
//   * emitted record = (k-p)/4 suffix bytes, most significant first, + counter bytes, least significant
//     first (:1198-1201); lut[prefix]++ (:1203) is aggregated per warp (match.any) and per tile (shared
//     window) before it touches global memory.
#pragma once
#include "common.cuh"
#include "radix_sort.cuh"

namespace kmcb {

struct CountArgs {
	const void* recs;        // sorted records
	uint64_t n;
	uint32_t n_tiles;
	uint32_t k;
	uint32_t lut_prefix_len;
	uint32_t cutoff_min, cutoff_max, counter_max;
	uint32_t counter_bytes;
	uint32_t suffix_bytes;   // (k - p) / 4
	uint8_t* out;            // emitted records
	uint64_t out_capacity;   // bytes
	uint64_t* lut;           // [4^p], zero-initialised
	uint64_t* result;        // [0..3] n_unique, n_cutoff_min, n_cutoff_max, n_total; [4] emitted records; [5] error (capacity)
	uint64_t* desc;          // [n_tiles] look-back chain
	uint32_t epoch;
	uint32_t* tile_counter;  // zero-initialised
	const uint32_t* run_flag; // nullptr: always run; else only when (*run_flag & kRunMask) == run_need (radix_sort.cuh)
	uint32_t run_need;
	const uint64_t* out_base; // nullptr or: records already emitted by earlier key blocks of an oversized bin (this launch appends behind them)
};

template <int WORDS> struct CountCfg { static constexpr int kThreads = 256, kIpt = (WORDS == 1 ? 8 : WORDS == 2 ? 4 : 2); };

constexpr int kLutWindow = 512;

template <int WORDS>
__host__ __device__ constexpr int count_tile() { return CountCfg<WORDS>::kThreads * CountCfg<WORDS>::kIpt; }

// dynamic shared memory layout: [records + 1 lookahead | staging of the emitted bytes | tail positions | emitted positions | emitted values]
template <int WORDS>
struct CountSmem {
	static constexpr int kTile = CountCfg<WORDS>::kThreads * CountCfg<WORDS>::kIpt;
	static constexpr size_t oRec = 0;
	static constexpr size_t oStage = ((size_t)(kTile + 1) * 8 * WORDS + 15) & ~(size_t)15;
	__host__ __device__ static size_t stage_bytes(uint32_t ob) { return ((size_t)kTile * ob + 32 + 15) & ~(size_t)15; }
	__host__ __device__ static size_t oTails(uint32_t ob) { return oStage + stage_bytes(ob); }
	__host__ __device__ static size_t oEmPos(uint32_t ob) { return oTails(ob) + (size_t)kTile * 2; }
	__host__ __device__ static size_t oEmVal(uint32_t ob) { return oEmPos(ob) + (size_t)kTile * 2; }
	__host__ __device__ static size_t bytes(uint32_t ob) { return oEmVal(ob) + (size_t)kTile * 4; }
};
template <int WORDS>
inline size_t count_smem_bytes(uint32_t out_rec_bytes) { return CountSmem<WORDS>::bytes(out_rec_bytes); }

// the p leading symbols of a k-mer (kmer.h:294-303 remove_suffix(2*(k-p)))
template <int WORDS>
__device__ __forceinline__ uint32_t rec_prefix(const Rec<WORDS>& r, uint32_t nbits)
{
	const uint32_t q = nbits >> 6, s = nbits & 63u;
	uint64_t lo = r.w[0], hi = 0;
#pragma unroll
	for (int i = 1; i < WORDS; ++i) {
		if (q == (uint32_t)i) lo = r.w[i];
		if (q + 1 == (uint32_t)i) hi = r.w[i];
	}
	if (WORDS == 1) return (uint32_t)(lo >> s);
	if (q + 1 >= (uint32_t)WORDS || s == 0) return (uint32_t)(lo >> s);
	return (uint32_t)((hi << (64u - s)) | (lo >> s));
}

// Sparse events (run tails, emitted records) are first compacted into dense lists, then handled one per thread:
// with 30x coverage only ~3 % of the records end a run, and a warp that carried the whole emit path for one
// active lane per round would spend ~30x the instructions.
template <int WORDS>
__global__ void __launch_bounds__(CountCfg<WORDS>::kThreads) count_emit_kernel(const CountArgs a)
{
	using R = Rec<WORDS>;
	using SM = CountSmem<WORDS>;
	constexpr int THREADS = CountCfg<WORDS>::kThreads, IPT = CountCfg<WORDS>::kIpt, TILE = THREADS * IPT;
	constexpr int WARPS = THREADS / 32, NW = TILE / 32;
	static_assert(NW <= THREADS, "one thread per bitmap word");
	extern __shared__ __align__(16) uint8_t dsm[];
	const uint32_t ob = a.suffix_bytes + a.counter_bytes;
	R* srec = reinterpret_cast<R*>(dsm + SM::oRec);
	uint8_t* stage0 = dsm + SM::oStage;
	uint16_t* tails_pos = reinterpret_cast<uint16_t*>(dsm + SM::oTails(ob));
	uint16_t* em_pos = reinterpret_cast<uint16_t*>(dsm + SM::oEmPos(ob));
	uint32_t* em_val = reinterpret_cast<uint32_t*>(dsm + SM::oEmVal(ob));
	__shared__ uint32_t tailmask[NW], wordpre[NW];
	__shared__ uint32_t lutwin[kLutWindow];
	__shared__ uint32_t s_tile, s_warp[WARPS], s_total;
	__shared__ uint64_t s_run_head0, s_base;

	if (!run_allowed(a.run_flag, a.run_need)) return;
	const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
	const R* __restrict__ g = reinterpret_cast<const R*>(a.recs);

	// persistent CTAs: tiles are claimed from a ticket (the chained scan needs them started in order)
	for (;;) {
	__syncthreads();
	if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1u);
	for (int i = tid; i < kLutWindow; i += THREADS) lutwin[i] = 0;
	__syncthreads();
	const uint32_t tile = s_tile;
	if (tile >= a.n_tiles) break;
	const uint64_t first = (uint64_t)tile * TILE;
	const uint64_t rem = a.n - first;
	const uint32_t cnt = rem < (uint64_t)TILE ? (uint32_t)rem : (uint32_t)TILE;

	// ---- load the tile (+ the records just before it for the run-head probe, + one lookahead record)
	R rec[IPT];
#pragma unroll
	for (int r = 0; r < IPT; ++r) {
		const uint32_t i = r * THREADS + tid;
		if (i < cnt) rec[r] = g[first + i];
	}
	R probe;
	const bool probe_valid = warp == 0 && first >= (uint64_t)lane + 1;
	if (probe_valid) probe = g[first - 1 - lane];
#pragma unroll
	for (int r = 0; r < IPT; ++r) {
		const uint32_t i = r * THREADS + tid;
		if (i < cnt) srec[i] = rec[r];
	}
	if (tid == 0 && first + cnt < a.n) srec[cnt] = g[first + cnt];
	__syncthreads();

	// ---- head of the run that contains the first record of the tile (warp 0): backward probe, then lower_bound for giant runs
	if (warp == 0) {
		uint64_t h = first;
		if (first != 0) {
			const R key0 = srec[0];
			uint64_t pos = first;
			bool found = false;
			for (int c = 0; c < 4 && !found; ++c) {
				bool neq = true;
				if (c == 0) { if (probe_valid) neq = !rec_equal<WORDS>(probe, key0); }
				else if (pos >= (uint64_t)lane + 1) neq = !rec_equal<WORDS>(g[pos - 1 - lane], key0);
				const uint32_t m = __ballot_sync(0xffffffffu, neq);
				if (m) {
					h = pos - (uint32_t)(__ffs(m) - 1);
					found = true;
				} else
					pos -= 32;
			}
			if (!found) {
				uint64_t lo = 0, hi = pos;
				while (lo < hi) {
					const uint64_t mid = (lo + hi) >> 1;
					if (rec_less<WORDS>(g[mid], key0)) lo = mid + 1;
					else hi = mid;
				}
				h = lo;
			}
		}
		if (lane == 0) s_run_head0 = h;
	}

	// ---- run tails: one neighbour compare per record, a bitmap word per warp round
	uint32_t tailbits = 0;
#pragma unroll
	for (int r = 0; r < IPT; ++r) {
		const uint32_t i = r * THREADS + tid;
		bool tail = false;
		if (i < cnt) tail = (first + i == a.n - 1) || !rec_equal<WORDS>(rec[r], srec[i + 1]);
		const uint32_t w = __ballot_sync(0xffffffffu, tail);
		if (lane == 0) tailmask[r * WARPS + warp] = w;
		tailbits |= (uint32_t)tail << r;
	}
	__syncthreads();

	// ---- exclusive scan of the bitmap popcounts -> dense list of tail positions
	{
		const uint32_t c = tid < NW ? __popc(tailmask[tid]) : 0;
		uint32_t inc = c;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t x = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= (uint32_t)o) inc += x;
		}
		if (lane == 31) s_warp[warp] = inc;
		__syncthreads();
		uint32_t base = 0;
		for (uint32_t w = 0; w < warp; ++w) base += s_warp[w];
		if (tid < NW) wordpre[tid] = base + inc - c;
		if (tid == NW - 1) s_total = base + inc;
	}
	__syncthreads();
	const uint32_t n_tails = s_total;
#pragma unroll
	for (int r = 0; r < IPT; ++r) {
		if ((tailbits >> r) & 1u) {
			const uint32_t i = r * THREADS + tid;
			const uint32_t wi = r * WARPS + warp;
			tails_pos[wordpre[wi] + __popc(tailmask[wi] & lanemask_lt())] = (uint16_t)i;
		}
	}
	__syncthreads();

	// ---- one thread per run: length, cutoffs (kb_sorter.h:1174-1191), dense list of the emitted ones
	uint32_t n_min = 0, n_max = 0, n_emit = 0;
	for (uint32_t j0 = 0; j0 < n_tails; j0 += THREADS) {
		const uint32_t j = j0 + tid;
		bool emit = false;
		uint32_t value = 0, pos = 0;
		bool is_min = false, is_max = false;
		if (j < n_tails) {
			pos = tails_pos[j];
			const uint64_t head = j ? first + tails_pos[j - 1] + 1 : s_run_head0;
			const uint32_t count = (uint32_t)(first + pos - head + 1);       // uint32 like kb_sorter.h:1153
			if (count < a.cutoff_min) is_min = true;
			else if (count > a.cutoff_max) is_max = true;
			else {
				emit = true;
				value = count > a.counter_max ? a.counter_max : count;
			}
		}
		const uint32_t be = __ballot_sync(0xffffffffu, emit);
		n_min += __popc(__ballot_sync(0xffffffffu, is_min));
		n_max += __popc(__ballot_sync(0xffffffffu, is_max));
		if (lane == 0) s_warp[warp] = __popc(be);
		__syncthreads();
		uint32_t base = n_emit, tot = 0;
#pragma unroll
		for (int w = 0; w < WARPS; ++w) {
			const uint32_t c = s_warp[w];
			if ((uint32_t)w < warp) base += c;
			tot += c;
		}
		if (emit) {
			const uint32_t e = base + __popc(be & lanemask_lt());
			em_pos[e] = (uint16_t)pos;
			em_val[e] = value;
		}
		n_emit += tot;
		__syncthreads();
	}
	const uint32_t emit_total = n_emit;
	if (tid == 0) {
		const uint64_t base = lookback_exclusive(a.desc, 1, tile, (uint64_t)emit_total, a.epoch);
		s_base = base;
		if (tile == a.n_tiles - 1) a.result[4] = base + emit_total;
	}
	if (lane == 0 && (n_min | n_max)) {      // every warp saw the same ballots only for its own lanes: per-warp partial sums
		if (n_min) atomicAdd(reinterpret_cast<unsigned long long*>(a.result) + 1, (unsigned long long)n_min);
		if (n_max) atomicAdd(reinterpret_cast<unsigned long long*>(a.result) + 2, (unsigned long long)n_max);
	}
	if (tid == 0) atomicAdd(reinterpret_cast<unsigned long long*>(a.result), (unsigned long long)n_tails);      // n_unique
	__syncthreads();

	// ---- one thread per emitted record: (k-p)/4 suffix bytes, most significant first, then the counter, least significant first
	const uint64_t base = s_base;
	const uint64_t ob0 = a.out_base ? *a.out_base : 0ull;
	uint8_t* dst = a.out + (ob0 + base) * ob;
	const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);     // staging keeps the destination's 16-byte phase
	uint8_t* stage = stage0 + mis;
	const bool fits = (ob0 + base + emit_total) * (uint64_t)ob <= a.out_capacity;
	const uint32_t prefix_shift = 2u * (a.k - a.lut_prefix_len);
	const uint32_t pfx0 = rec_prefix<WORDS>(srec[0], prefix_shift);
	for (uint32_t e0 = 0; e0 < emit_total; e0 += THREADS) {
		const uint32_t e = e0 + tid;
		uint32_t pfx = 0xffffffffu;
		if (e < emit_total) {
			const R rr = srec[em_pos[e]];
			const uint32_t value = em_val[e];
			uint8_t* o = stage + (size_t)e * ob;
			for (uint32_t j = 0; j < a.suffix_bytes; ++j) o[j] = (uint8_t)rec_byte<WORDS>(rr, a.suffix_bytes - 1 - j);
			for (uint32_t j = 0; j < a.counter_bytes; ++j) o[a.suffix_bytes + j] = (uint8_t)(value >> (8 * j));
			pfx = rec_prefix<WORDS>(rr, prefix_shift);
		}
		// lut[prefix]++ (kb_sorter.h:1203) aggregated per warp, then per tile in a shared window
		const uint32_t peers = __match_any_sync(0xffffffffu, pfx);
		if (e < emit_total && (int)lane == __ffs(peers) - 1) {
			const uint32_t c = __popc(peers);
			const uint32_t d = pfx - pfx0;
			if (d < (uint32_t)kLutWindow) atomicAdd(&lutwin[d], c);
			else atomicAdd(reinterpret_cast<unsigned long long*>(a.lut) + pfx, (unsigned long long)c);
		}
	}
	__syncthreads();

	// ---- copy the staged bytes out: head bytes up to 16-byte alignment, 16-byte vectors, tail bytes
	if (fits) {
		const uint32_t total = emit_total * ob;
		const uint32_t headb = min(total, (16u - mis) & 15u);
		if (tid < headb) dst[tid] = stage[tid];
		const uint32_t nvec = (total - headb) >> 4;
		const uint4* sv = reinterpret_cast<const uint4*>(stage + headb);
		uint4* dv = reinterpret_cast<uint4*>(dst + headb);
		for (uint32_t v = tid; v < nvec; v += THREADS) dv[v] = sv[v];
		const uint32_t done = headb + (nvec << 4);
		if (tid < total - done) dst[done + tid] = stage[done + tid];
	} else if (tid == 0)
		a.result[5] = 1;

	for (int i = tid; i < kLutWindow; i += THREADS) {
		const uint32_t c = lutwin[i];
		if (c) atomicAdd(reinterpret_cast<unsigned long long*>(a.lut) + pfx0 + i, (unsigned long long)c);
	}
	}      // tiles
}

}  // namespace kmcb
