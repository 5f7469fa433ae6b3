// kmc_b200 — leaves of the hybrid MSD path for one-word k-mers (k <= 32): COUNT WITHOUT SORTING THE DUPLICATES.
//
// What stage 2 needs from a leaf bucket is the sorted list of its DISTINCT k-mers with their multiplicities
// (CompactKmers, kmc_core/kb_sorter.h:1128-1281).  With sequencing coverage c a k-mer occurs ~c times, so sorting every
// copy (RADULS, or our own leaf sort) does ~c times the necessary work.  A leaf (records that share their top 8+b2 bits,
// ~1 K records) is therefore counted in an order-preserving table in shared memory instead:
//   * slot = the next 11 bits of the k-mer, claimed with one 64-bit atomicCAS; a copy of a k-mer already there is one
//     atomicAdd.  Slot order is key order, so no sort is needed;
//   * two different k-mers in one slot (same next 11 bits) are rare (~15 % of the distinct k-mers at 30x coverage): the later
//     one goes to a small open-addressing side table; side entries are ranked against the few entries of their own slot;
//   * cutoffs / clamping / record bytes / LUT exactly as kb_sorter.h:1174-1203; a prefix sum over the slots gives every
//     surviving k-mer its position, the records are staged in shared memory and written to the leaf's region of a
//     temporary buffer; a tiny scan over the leaves + leaf_gather_kernel packs the regions into the output.
// No look-back, no inter-CTA dependency.  A leaf whose side table overflows (heavy skew) raises the fallback flag:
// the LSD passes + count_emit_kernel then redo the bin (msd_sort.cuh).
#pragma once
#include "common.cuh"
#include "msd_sort.cuh"

namespace kmcb {

constexpr int kLeafThreads = 256;
constexpr int kLeafSlotBits = 11;
constexpr int kLeafSlots = 1 << kLeafSlotBits;      // main table
constexpr int kLeafSide = 512;                      // side table (open addressing)
constexpr int kLeafSideMax = 448;
constexpr uint64_t kLeafEmpty = ~0ull;

struct LeafArgs {
	const uint64_t* recs;        // partitioned records
	const uint64_t* start;       // [n_leaves + 1]
	uint32_t n_leaves;
	uint32_t low_bits;           // bits below the partition digits
	uint32_t k, lut_prefix_len, cutoff_min, cutoff_max, counter_max, counter_bytes, suffix_bytes;
	uint8_t* tmp;                // leaf L writes its records, padded to 8 (16) bytes, at tmp + start[L] * 8 (16)
	uint32_t* leaf_emit;         // [n_leaves] emitted records
	uint64_t* lut;
	uint64_t* result;            // [0] n_unique [1] n_cutoff_min [2] n_cutoff_max
	uint32_t* ticket;
	uint32_t* flags;
};

constexpr int kLeafRoundRecords = 2560;              // records one round of a leaf is sized for
constexpr int kLeafRetry = 2048;                    // copies of k-mers that lost their slot to another k-mer, handled in a dense second round
constexpr int kLeafWords = kLeafSlots / 32;

// Everything after the counting works on DENSE lists (occupied main slots, side entries, survivors): only ~15 % of the slots
// are occupied and ~3 % of the records end up in the output, and a warp that sweeps all slots or carries a rare path for one
// active lane pays the full instruction count anyway.
struct LeafSmem {
	uint64_t mkey[kLeafSlots];       // 16 KB   main table: k-mers
	uint32_t mcnt[kLeafSlots];       //  8 KB   main table: multiplicities
	uint16_t occ[kLeafSlots];        //  4 KB   occupied main slots, in order of arrival
	uint32_t surv[kLeafWords];       //         bitmap: the slot's main entry survives the cutoffs
	uint32_t sidew[kLeafWords];      //         surviving side entries per 32 slots
	uint32_t wpre[kLeafWords];       //         exclusive prefix of popc(surv) + sidew = position of the word's first survivor
	uint64_t skey[kLeafSide];        //  4 KB   side table (open addressing)
	uint32_t scnt[kLeafSide];        //  2 KB
	uint16_t sslot[kLeafSide];       //  1 KB   main slot the side entry belongs to
	uint16_t socc[kLeafSide];        //  1 KB   occupied side slots
	uint16_t dense[kLeafSide];       //  1 KB   surviving side entries
	uint32_t retry[kLeafRetry];      //  8 KB   indices (inside the leaf) of records whose slot was taken
	uint32_t n_occ, n_side, n_dense, n_allones, n_retry, total_emit;
};

// record image: (k-p)/4 suffix bytes most significant first, then the counter least significant first (kb_sorter.h:1198-1201),
// as a little-endian integer (byte 0 of the record = bits 0-7)
__device__ __forceinline__ void leaf_record(uint64_t key, uint32_t value, const LeafArgs& a, uint64_t& lo, uint32_t& hi)
{
	const uint32_t sb = a.suffix_bytes;
	uint64_t suf = sb >= 8 ? key : (key & ((1ull << (8 * sb)) - 1));
	suf = (uint64_t)__byte_perm((uint32_t)(suf >> 32), 0, 0x0123) | ((uint64_t)__byte_perm((uint32_t)suf, 0, 0x0123) << 32);     // byte-reversed: most significant byte first in memory ...
	suf = sb >= 8 ? suf : (suf >> (8 * (8 - sb)));                                                                              // ... starting at byte 0
	lo = suf;
	hi = 0;
	if (sb < 8) lo |= (uint64_t)value << (8 * sb);
	if (sb + a.counter_bytes > 8) hi = sb >= 8 ? value : (value >> (8 * (8 - sb)));
}

// cutoffs and clamp (kb_sorter.h:1174-1191): 0 = below cutoff_min, 1 = above cutoff_max, 2 = emitted
__device__ __forceinline__ uint32_t leaf_class(uint32_t c, const LeafArgs& a) { return c < a.cutoff_min ? 0u : c > a.cutoff_max ? 1u : 2u; }

__global__ void __launch_bounds__(kLeafThreads, 4) leaf_count_kernel(const LeafArgs a)
{
	extern __shared__ __align__(16) uint8_t dsm[];
	LeafSmem& S = *reinterpret_cast<LeafSmem*>(dsm);
	if (*a.flags & kMsdFlagFallback) return;
	const uint32_t tid = threadIdx.x, lane = tid & 31u;
	const uint32_t pad8 = (a.suffix_bytes + a.counter_bytes) > 8 ? 2u : 1u;          // temporary records: 8 or 16 bytes
	const uint32_t prefix_shift = 2u * (a.k - a.lut_prefix_len);
	const bool one_prefix = prefix_shift >= a.low_bits;        // every k-mer of a leaf has the same LUT prefix
	const bool maybe_allones = a.k == 32;
	uint32_t n_unique = 0, n_min = 0, n_max = 0;
	bool failed = false;
	uint64_t* const tmp64 = reinterpret_cast<uint64_t*>(a.tmp);

	// Leaves are dealt round-robin (leaf = blockIdx.x, + gridDim.x, ...): every CTA gets ~100 of them, so sizes average out, and
	// knowing the next leaf in advance lets its boundaries be loaded and its records be pulled into L2 while the current one is counted.
	uint32_t leaf = blockIdx.x;
	uint64_t lo = 0, hi = 0;
	if (leaf < a.n_leaves) { lo = a.start[leaf]; hi = a.start[leaf + 1]; }
	for (; leaf < a.n_leaves; leaf += gridDim.x) {
		const uint32_t m = (uint32_t)(hi - lo);
		const uint64_t cur_lo = lo;
		{
			const uint32_t nl = leaf + gridDim.x;
			if (nl < a.n_leaves) { lo = a.start[nl]; hi = a.start[nl + 1]; }
		}
		if (m == 0) { if (tid == 0) a.leaf_emit[leaf] = 0; continue; }
		// A leaf with more records than the tables are made for (canonical k-mers crowd into the low prefixes: up to ~4x the average)
		// is counted in 2^e rounds: round r takes the k-mers whose next e bits are r, so the rounds' outputs simply follow each other.
		uint32_t e_bits = 0;
		while ((m >> e_bits) > (uint32_t)kLeafRoundRecords && e_bits < 8 && e_bits < a.low_bits) ++e_bits;
		if ((m >> e_bits) > (uint32_t)kLeafRoundRecords * 2) failed = true;
		const uint32_t sub_shift = a.low_bits - e_bits;
		const uint32_t slot_shift = sub_shift > (uint32_t)kLeafSlotBits ? sub_shift - kLeafSlotBits : 0;
		uint32_t emit_base = 0;
		for (uint32_t round = 0; round < (1u << e_bits); ++round) {
			__syncthreads();          // the previous leaf / round is completely done with the tables
			// ---- clear (16-byte stores)
			{
				uint4* k4 = reinterpret_cast<uint4*>(S.mkey);
				const uint4 e = make_uint4(~0u, ~0u, ~0u, ~0u), z = make_uint4(0, 0, 0, 0);
#pragma unroll
				for (int i = 0; i < kLeafSlots * 8 / 16 / kLeafThreads; ++i) k4[i * kLeafThreads + tid] = e;
				uint4* c4 = reinterpret_cast<uint4*>(S.mcnt);
#pragma unroll
				for (int i = 0; i < kLeafSlots * 4 / 16 / kLeafThreads; ++i) c4[i * kLeafThreads + tid] = z;
				reinterpret_cast<uint4*>(S.skey)[tid] = e;                                  // 512 * 8 B = 256 * 16 B
				reinterpret_cast<uint2*>(S.scnt)[tid] = make_uint2(0, 0);
				if (tid < kLeafWords) { S.surv[tid] = 0; S.sidew[tid] = 0; }
				if (tid == 0) { S.n_occ = 0; S.n_side = 0; S.n_dense = 0; S.n_allones = 0; S.n_retry = 0; }
			}
			__syncthreads();

			// ---- count, round 1: slot = next 11 bits of the k-mer.  First arrival claims the slot (and notes it in the occupied list),
			// copies add one, a record whose slot belongs to another k-mer is only noted down.
			for (uint32_t j0 = 0; j0 < m; j0 += 4 * kLeafThreads) {
				uint64_t key[4];
#pragma unroll
				for (int u = 0; u < 4; ++u) { const uint32_t j = j0 + u * kLeafThreads + tid; key[u] = j < m ? a.recs[cur_lo + j] : 0; }
				// all four claims are issued before any result is looked at: their latencies overlap
				uint32_t slot[4];
				unsigned long long old[4];
				uint32_t live = 0;
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					const uint32_t j = j0 + u * kLeafThreads + tid;
					const uint64_t kk = key[u];
					bool ok = j < m;
					if (e_bits && ((uint32_t)(kk >> sub_shift) & ((1u << e_bits) - 1u)) != round) ok = false;      // another round's k-mer
					if (ok && maybe_allones && kk == kLeafEmpty) { atomicAdd(&S.n_allones, 1u); ok = false; }       // TTT..T (k = 32, -b): cannot live in the table, sorts last
					slot[u] = (uint32_t)(kk >> slot_shift) & (kLeafSlots - 1);
					old[u] = kk;
					if (ok) {
						old[u] = atomicCAS(reinterpret_cast<unsigned long long*>(&S.mkey[slot[u]]), (unsigned long long)kLeafEmpty, (unsigned long long)kk);
						live |= 1u << u;
					}
				}
#pragma unroll
				for (int u = 0; u < 4; ++u) {
					if (!((live >> u) & 1u)) continue;
					const uint64_t kk = key[u];
					const uint32_t b = slot[u];
					if (old[u] == kLeafEmpty) S.occ[atomicAdd(&S.n_occ, 1u)] = (uint16_t)b;
					if (old[u] == kLeafEmpty || old[u] == kk) atomicAdd(&S.mcnt[b], 1u);
					else {
						const uint32_t q = atomicAdd(&S.n_retry, 1u);
						if (q < (uint32_t)kLeafRetry) S.retry[q] = j0 + u * kLeafThreads + tid;
						else failed = true;
					}
				}
			}
			if (round == 0 && leaf + gridDim.x < a.n_leaves) {        // pull the next leaf towards L2 (128 bytes per thread and step)
				const uint32_t nm = (uint32_t)(hi - lo);
				for (uint32_t i = tid * 16; i < nm; i += kLeafThreads * 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.recs + lo + i));
			}
			__syncthreads();
			// ---- count, round 2 (dense): the noted records go to the side table
			{
				const uint32_t nr = min(S.n_retry, (uint32_t)kLeafRetry);
				for (uint32_t q = tid; q < nr; q += kLeafThreads) {
					const uint64_t kk = a.recs[cur_lo + S.retry[q]];
					const uint32_t b = (uint32_t)(kk >> slot_shift) & (kLeafSlots - 1);
					uint32_t h = (uint32_t)((kk * 0x9E3779B97F4A7C15ull) >> 55) & (kLeafSide - 1);
					int probe = 0;
					for (; probe < kLeafSide; ++probe) {
						const unsigned long long o2 = atomicCAS(reinterpret_cast<unsigned long long*>(&S.skey[h]), (unsigned long long)kLeafEmpty, (unsigned long long)kk);
						if (o2 == kLeafEmpty) {
							S.sslot[h] = (uint16_t)b;
							const uint32_t i = atomicAdd(&S.n_side, 1u);
							if (i < (uint32_t)kLeafSide) S.socc[i] = (uint16_t)h;
						}
						if (o2 == kLeafEmpty || o2 == kk) { atomicAdd(&S.scnt[h], 1u); break; }
						h = (h + 1) & (kLeafSide - 1);
					}
					if (probe == kLeafSide) failed = true;
				}
			}
			__syncthreads();
			const uint32_t n_occ = S.n_occ, n_side = min(S.n_side, (uint32_t)kLeafSide);
			if (S.n_side > (uint32_t)kLeafSideMax) failed = true;

			// ---- cutoffs (dense): survivors of the main table as a bitmap, surviving side entries as a list + a count per 32 slots
			for (uint32_t q = tid; q < n_occ; q += kLeafThreads) {
				const uint32_t b = S.occ[q];
				const uint32_t cl = leaf_class(S.mcnt[b], a);
				n_min += cl == 0;
				n_max += cl == 1;
				if (cl == 2) atomicOr(&S.surv[b >> 5], 1u << (b & 31));
			}
			for (uint32_t q = tid; q < n_side; q += kLeafThreads) {
				const uint32_t h = S.socc[q];
				const uint32_t cl = leaf_class(S.scnt[h], a);
				n_min += cl == 0;
				n_max += cl == 1;
				if (cl == 2) {
					S.dense[atomicAdd(&S.n_dense, 1u)] = (uint16_t)h;
					atomicAdd(&S.sidew[S.sslot[h] >> 5], 1u);
				}
			}
			if (tid == 0) {
				n_unique += n_occ + n_side;
				if (S.n_allones) { const uint32_t cl = leaf_class(S.n_allones, a); ++n_unique; n_min += cl == 0; n_max += cl == 1; }
			}
			__syncthreads();

			// ---- exclusive prefix over the 64 bitmap words (two warps)
			if (tid < kLeafWords) {
				const uint32_t v = __popc(S.surv[tid]) + S.sidew[tid];
				uint32_t inc = v;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
					if (lane >= (uint32_t)o) inc += t;
				}
				S.wpre[tid] = inc - v;                       // per-warp exclusive prefix; the second warp adds the first warp's total below
				if (tid == 31) S.total_emit = inc;
				__syncwarp();
			}
			__syncthreads();
			if (tid >= 32 && tid < kLeafWords) S.wpre[tid] += S.total_emit;
			__syncthreads();
			const uint32_t n_dense = S.n_dense;
			const uint32_t total = S.wpre[kLeafWords - 1] + __popc(S.surv[kLeafWords - 1]) + S.sidew[kLeafWords - 1];
			const uint32_t allones_emit = (S.n_allones && leaf_class(S.n_allones, a) == 2) ? 1u : 0u;

			// ---- emission (dense): every survivor computes its own position and stores its record (one aligned 8-byte store)
			auto emit = [&](uint32_t pos, uint64_t kk, uint32_t c) {
				const uint32_t value = c > a.counter_max ? a.counter_max : c;          // kb_sorter.h:1190
				uint64_t rl; uint32_t rh;
				leaf_record(kk, value, a, rl, rh);
				tmp64[(cur_lo + emit_base + pos) * pad8] = rl;
				if (pad8 == 2) tmp64[(cur_lo + emit_base + pos) * 2 + 1] = rh;
				if (!one_prefix) atomicAdd(reinterpret_cast<unsigned long long*>(a.lut) + (kk >> prefix_shift), 1ull);     // kb_sorter.h:1203
			};
			if (!failed) {
				for (uint32_t q = tid; q < n_occ; q += kLeafThreads) {
					const uint32_t b = S.occ[q], w = b >> 5;
					const uint32_t bits = S.surv[w];
					if (!((bits >> (b & 31)) & 1u)) continue;
					const uint64_t kk = S.mkey[b];
					uint32_t pos = S.wpre[w] + __popc(bits & ((1u << (b & 31)) - 1u));
					if (S.sidew[w])                               // surviving side entries in this word: those that sort before this k-mer
						for (uint32_t f = 0; f < n_dense; ++f) {
							const uint32_t h2 = S.dense[f], b2 = S.sslot[h2];
							if ((b2 >> 5) == w && (b2 < b || (b2 == b && S.skey[h2] < kk))) ++pos;
						}
					emit(pos, kk, S.mcnt[b]);
				}
				for (uint32_t e = tid; e < n_dense; e += kLeafThreads) {
					const uint32_t h = S.dense[e], b = S.sslot[h], w = b >> 5;
					const uint64_t kk = S.skey[h];
					const uint32_t bits = S.surv[w];
					uint32_t pos = S.wpre[w] + __popc(bits & ((1u << (b & 31)) - 1u));
					if (((bits >> (b & 31)) & 1u) && S.mkey[b] < kk) ++pos;          // the main entry of the own slot
					if (S.sidew[w] > 1)
						for (uint32_t f = 0; f < n_dense; ++f) {
							const uint32_t h2 = S.dense[f], b2 = S.sslot[h2];
							if ((b2 >> 5) == w && (b2 < b || (b2 == b && S.skey[h2] < kk))) ++pos;
						}
					emit(pos, kk, S.scnt[h]);
				}
				if (tid == 0 && allones_emit) emit(total, kLeafEmpty, S.n_allones);
			}
			emit_base += total + allones_emit;
		}      // rounds
		if (tid == 0) {
			a.leaf_emit[leaf] = emit_base;
			if (one_prefix && emit_base)
				atomicAdd(reinterpret_cast<unsigned long long*>(a.lut) + (((uint64_t)leaf << a.low_bits) >> prefix_shift), (unsigned long long)emit_base);
		}
	}
	// ---- statistics of this CTA
	if (failed) atomicOr(a.flags, kMsdFlagFallback);
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		n_unique += __shfl_down_sync(0xffffffffu, n_unique, o);
		n_min += __shfl_down_sync(0xffffffffu, n_min, o);
		n_max += __shfl_down_sync(0xffffffffu, n_max, o);
	}
	if (lane == 0) {
		if (n_unique) atomicAdd(reinterpret_cast<unsigned long long*>(a.result), (unsigned long long)n_unique);
		if (n_min) atomicAdd(reinterpret_cast<unsigned long long*>(a.result) + 1, (unsigned long long)n_min);
		if (n_max) atomicAdd(reinterpret_cast<unsigned long long*>(a.result) + 2, (unsigned long long)n_max);
	}
}

// exclusive scan of the per-leaf record counts (single CTA, every thread owns up to 64 consecutive leaves), total -> result[4]
__global__ void __launch_bounds__(1024) leaf_scan_kernel(const uint32_t* leaf_emit, uint32_t n_leaves, uint64_t* leaf_off, uint64_t* result, uint64_t out_capacity, uint32_t ob, const uint32_t* flags)
{
	__shared__ uint64_t s_w[32];
	if (*flags & kMsdFlagFallback) return;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t per = (n_leaves + 1023) / 1024;          // <= 64
	const uint32_t i0 = tid * per;
	uint64_t sum = 0;
	for (uint32_t i = 0; i < per; ++i) sum += (i0 + i < n_leaves) ? leaf_emit[i0 + i] : 0u;
	uint64_t inc = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= (uint32_t)o) inc += t;
	}
	if (lane == 31) s_w[warp] = inc;
	__syncthreads();
	uint64_t base = inc - sum, tot = 0;
	for (uint32_t w = 0; w < 32; ++w) { if (w < warp) base += s_w[w]; tot += s_w[w]; }
	for (uint32_t i = 0; i < per; ++i) {
		if (i0 + i < n_leaves) { leaf_off[i0 + i] = base; base += leaf_emit[i0 + i]; }
	}
	if (tid == 0) {
		result[4] = tot;
		if (tot * ob > out_capacity) result[5] = 1;
	}
}

// one warp per leaf: its padded temporary records -> packed records at their final place
__global__ void __launch_bounds__(256) leaf_gather_kernel(const uint8_t* tmp, const uint64_t* start, const uint32_t* leaf_emit, const uint64_t* leaf_off,
	uint32_t n_leaves, uint32_t ob, uint8_t* out, const uint64_t* result, const uint32_t* flags)
{
	if (*flags & kMsdFlagFallback) return;
	if (result[5]) return;                        // capacity error: nothing is written
	const uint32_t leaf = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31u;
	if (leaf >= n_leaves) return;
	const uint32_t pad = ob > 8 ? 16u : 8u;
	const uint32_t nbytes = leaf_emit[leaf] * ob;
	const uint8_t* src = tmp + start[leaf] * pad;
	uint8_t* dst = out + leaf_off[leaf] * ob;
	const uint32_t magic = 0xFFFFFFFFu / ob + 1;          // p / ob == umulhi(p, magic) for p < 2^16 ... checked: larger leaves take the division
	for (uint32_t p = lane; p < nbytes; p += 32) {
		const uint32_t r = nbytes < 65536u ? __umulhi(p, magic) : p / ob;
		dst[p] = src[(size_t)r * pad + (p - r * ob)];
	}
}

// when the hybrid path gave up after the leaves had already touched lut / result: start over for the fallback
__global__ void leaf_reset_kernel(uint64_t* lut, uint64_t lut_entries, uint64_t* result, const uint32_t* flags)
{
	if (!(*flags & kMsdFlagFallback)) return;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < lut_entries; i += (uint64_t)gridDim.x * blockDim.x) lut[i] = 0;
	if (blockIdx.x == 0 && threadIdx.x < 6) result[threadIdx.x] = 0;
}

}  // namespace kmcb
