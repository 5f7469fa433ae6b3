// TEST INFRASTRUCTURE — not part of the product.
//
// Harness around the UNMODIFIED reference KMC stage-2 classes.  It is compiled by
// oracle/Makefile against the sources where they lie under /root/reference (nothing
// is copied) into oracle/_ref/libkmc_ref.so and exposes a tiny C ABI so that
// tests/ and bench.py (cpu_baseline / --impl reference) can run
//   * O2: the reference's own per-bin stage 2, CKmerBinSorter<SIZE>::ProcessBins
//         (kmc_core/kb_sorter.h:210-237) fed through the real CBinDesc / CExpanderPackDesc /
//         CMemoryBins / CBinQueue / CSortersManager / CKmerQueue objects exactly as
//         CKmerBinReader::ProcessBins does (kmc_core/kb_reader.h:104-224) and drained
//         exactly as CKmerBinCompleter::ProcessBinsFirstStage does (kb_completer.cpp:131-205);
//   * O3: the reference's sort_func alone, RadulsSort::RadixSortMSD_AVX2<CKmer<SIZE>>
//         (kmc_core/raduls_impl.h:769-776) or RadixSort::RadixSortMSD (kmc_core/radix.h:845-855).
//
// Only the harness glue below is ours; every algorithmic step runs reference code.

#include "kmc_core/defs.h"
#include "kmc_core/params.h"
#include "kmc_core/kmer.h"
#include "kmc_core/raduls.h"
#include "kmc_core/radix.h"
#include "kmc_core/kb_sorter.h"
#ifdef KMCREF_WITH_B200
// drop-in test: the product's host shim compiled inside the reference tree, in place of CKmerBinSorter
#include "kb_sorter_b200.h"
#endif

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <functional>
#include <atomic>

namespace {

struct BinIO {
	const uint8_t* data; uint64_t size; uint64_t n_rec; uint64_t n_plus_x_recs;
	const uint64_t* pack_bytes; const uint64_t* pack_recs; uint32_t n_packs;
	uint8_t* out; uint64_t out_cap; uint64_t out_bytes;
	uint64_t* lut; uint64_t stats[4];
};

inline int64_t RU(int64_t x) { return (x + ALIGNMENT - 1) / ALIGNMENT * ALIGNMENT; }

double now_s() {
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <unsigned SIZE>
int run_bins(int k, int both_strands, uint32_t cutoff_min, uint32_t cutoff_max, uint32_t counter_max,
	uint32_t lut_prefix_len, int n_sorters, int sort_kind, std::vector<BinIO>& bins, double* times)
{
	// sort_kind: 0 RADULS, 1 radix.h, 2 = the B200 drop-in (CKmerBinSorterB200 instead of CKmerBinSorter)
	const int n_bins = (int)bins.size();
	CKMCParams P{};
	P.kmer_len = k;
	P.both_strands = both_strands != 0;
	P.n_bins = n_bins;
	P.cutoff_min = (int)cutoff_min;
	P.cutoff_max = cutoff_max;
	P.counter_max = counter_max;
	P.max_x = (k % 32 == 0) ? 0 : MIN(31 - (k % 32), KMER_X);       // kmc_core/kmc.h:139-142
	P.without_output = false;
	P.lut_prefix_len = lut_prefix_len;
	P.output_type = OutputType::KMC;
	P.n_sorters = n_sorters;
	P.n_threads = n_sorters;
	P.use_strict_mem = false;

	CKMCQueues Q;
	Q.bd = std::make_unique<CBinDesc>(k, n_bins);
	Q.epd = std::make_unique<CExpanderPackDesc>(n_bins);
	Q.bq = std::make_unique<CBinQueue>(1);
	Q.kq = std::make_unique<CKmerQueue>(n_bins, n_sorters);

	// kmc_core/kmc.h:370-379
	int64_t part = (256 * GetBufferWidth(sizeof(CKmer<SIZE>) / 8) + ALIGNMENT) * sizeof(CKmer<SIZE>);
	Q.pmm_radix_buf = std::make_unique<CMemoryPool>(part * n_sorters * MAGIC_NUMBER, part);

	for (int b = 0; b < n_bins; ++b) {
		Q.bd->insert(b, nullptr, "bin");
		Q.bd->update(b, (int64_t)bins[b].size, bins[b].n_rec, bins[b].n_plus_x_recs, 0);
		std::list<std::pair<uint64, uint64>> l;
		for (uint32_t i = 0; i < bins[b].n_packs; ++i)
			l.emplace_back(bins[b].pack_bytes[i], bins[b].pack_recs[i]);
		Q.epd->push(b, l);
	}
	auto sorted = Q.bd->get_sorted_req_sizes(P.max_x, sizeof(CKmer<SIZE>), P.cutoff_min, P.cutoff_max, P.counter_max, P.lut_prefix_len);
	Q.bd->init_sort(sorted);

	// arena: like kmc.h:1500-1572 — sum of requirements, never smaller than the largest bin
	int64_t total = 0;
	for (auto& s : sorted) total += s.second;
	int64_t max_mem = MAX(total, (int64_t)16 << 20);
	// keep the harness inside this box's RAM: KMCREF_ARENA_GB (bench.py sets it from MemAvailable and sweeps it - the arena size is
	// what decides how many bins the reference works on at the same time, like kmc's -m), default 48 GB
	int64_t cap = (int64_t)48 << 30;
	if (const char* e = getenv("KMCREF_ARENA_GB")) { long long g = atoll(e); if (g >= 1) cap = (int64_t)g << 30; }
	if (max_mem > cap) max_mem = MAX(cap, sorted.front().second);
	// The arena is kept across calls (a real stage 2 allocates it once for all its bins, kmc.h:1510, so later bins run on
	// pages that are already faulted in); the warm-up steps of bench.py play the role of the earlier bins.
	static std::unique_ptr<CMemoryBins> arena_cache;
	static int64_t arena_key[3] = { -1, -1, -1 };
	if (arena_cache && arena_key[0] == max_mem && arena_key[1] == n_bins && arena_key[2] == n_sorters)
		Q.memory_bins = std::move(arena_cache);
	else {
		arena_cache.reset();
		Q.memory_bins = std::make_unique<CMemoryBins>(max_mem + (1 << 20), n_bins, false, n_sorters);
	}
	arena_key[0] = max_mem; arena_key[1] = n_bins; arena_key[2] = n_sorters;
	int64_t mm = Q.memory_bins->GetTotalSize();
	if (mm < sorted.front().second) mm = sorted.front().second;
	Q.sorters_manager = std::make_unique<CSortersManager>(n_bins, n_sorters, Q.bq.get(), mm, sorted);

	double t_sort_acc = 0;
	std::mutex t_mtx;
	SortFunction<CKmer<SIZE>> base;
	if (sort_kind == 0) base = RadulsSort::RadixSortMSD_AVX2<CKmer<SIZE>>;
	else { base = RadixSort::RadixSortMSD<CKmer<SIZE>, SIZE>; CSmallSort<SIZE>::Adjust(384); }
	SortFunction<CKmer<SIZE>> sort_func = [&](CKmer<SIZE>* a, CKmer<SIZE>* t, uint64 n, uint32 byte, uint32 thr, CMemoryPool* pool) {
		double t0 = now_s();
		base(a, t, n, byte, thr, pool);
		double t1 = now_s();
		std::lock_guard<std::mutex> l(t_mtx);
		t_sort_acc += t1 - t0;
	};

	double t_begin = now_s();

	// sorter threads (kmc.h:1576-1584)
	std::vector<std::unique_ptr<CKmerBinSorter<SIZE>>> sorters;
	std::vector<std::thread> sorter_threads;
#ifdef KMCREF_WITH_B200
	std::vector<std::unique_ptr<CKmerBinSorterB200<SIZE>>> gpu_sorters;
	if (sort_kind == 2) {
		for (int i = 0; i < n_sorters; ++i)
			gpu_sorters.emplace_back(std::make_unique<CKmerBinSorterB200<SIZE>>(P, Q, 0));
		for (int i = 0; i < n_sorters; ++i)
			sorter_threads.emplace_back([&, i] { gpu_sorters[i]->ProcessBins(); });
	} else
#endif
	{
		if (sort_kind == 2) return -4;
		for (int i = 0; i < n_sorters; ++i)
			sorters.emplace_back(std::make_unique<CKmerBinSorter<SIZE>>(P, Q, sort_func));
		for (int i = 0; i < n_sorters; ++i)
			sorter_threads.emplace_back([&, i] { sorters[i]->ProcessBins(); });
	}

	// completer stand-in (kb_completer.cpp:131-205): copy the packs, free suffix + lut
	std::atomic<int> rc{0};
	std::thread completer([&] {
		int32 bin_id; uchar* data; std::list<std::pair<uint64, uint64>> packs; uchar* lut; uint64 lut_size;
		uint64 nu, ncmin, ncmax, nt;
		while (Q.kq->pop(bin_id, data, packs, lut, lut_size, nu, ncmin, ncmax, nt)) {
			BinIO& B = bins[bin_id];
			uint64_t pos = 0;
			for (auto& p : packs) {
				uint64_t len = p.second - p.first;
				if (pos + len > B.out_cap) { rc = -2; break; }
				memcpy(B.out + pos, data + p.first, len);
				pos += len;
			}
			B.out_bytes = pos;
			if (B.lut && lut_size) memcpy(B.lut, lut, lut_size);
			B.stats[0] = nu; B.stats[1] = ncmin; B.stats[2] = ncmax; B.stats[3] = nt;
			Q.memory_bins->free(bin_id, CMemoryBins::mba_suffix);
			Q.memory_bins->free(bin_id, CMemoryBins::mba_lut);
		}
	});

	// reader stand-in (kb_reader.h:120-196)
	int32 bin_id;
	while ((bin_id = Q.bd->get_next_sort_bin()) >= 0) {
		BinIO& B = bins[bin_id];
		uint64 input_kmer_size, kxmer_counter_size; uint32 kxmer_symbols;
		if (P.max_x) {
			input_kmer_size = B.n_plus_x_recs * sizeof(CKmer<SIZE>);
			kxmer_counter_size = B.n_plus_x_recs * sizeof(uint32);
			kxmer_symbols = k + P.max_x + 1;
		} else {
			input_kmer_size = B.n_rec * sizeof(CKmer<SIZE>);
			kxmer_counter_size = 0;
			kxmer_symbols = k;
		}
		uint64 max_out_recs = (B.n_rec + 1) / max(cutoff_min, 1u);
		uint64 counter_size = calc_counter_size(cutoff_max, counter_max);
		uint32 kmer_symbols = k - lut_prefix_len;
		uint64 kmer_bytes = kmer_symbols / 4;
		uint64 out_buffer_size = max_out_recs * (kmer_bytes + counter_size);
		uint32 rec_len = (kxmer_symbols + 3) / 4;
		uint64 lut_size = (1ull << (2 * lut_prefix_len)) * sizeof(uint64);

		Q.memory_bins->init(bin_id, rec_len, RU(B.size), RU(input_kmer_size), RU(out_buffer_size), RU(kxmer_counter_size), RU(lut_size));
		uchar* data;
		if (B.size > 0) {
			Q.memory_bins->reserve(bin_id, data, CMemoryBins::mba_input_file);
			memcpy(data, B.data, B.size);
			Q.memory_bins->extend(bin_id, rec_len, RU(B.size), RU(input_kmer_size), RU(out_buffer_size), RU(kxmer_counter_size), RU(lut_size));
			Q.memory_bins->reserve(bin_id, data, CMemoryBins::mba_input_file);
			Q.bq->push(bin_id, data, B.size, B.n_rec);
		} else {
			Q.memory_bins->extend(bin_id, rec_len, RU(B.size), RU(input_kmer_size), RU(out_buffer_size), RU(kxmer_counter_size), RU(lut_size));
			Q.bq->push(bin_id, nullptr, 0, 0);
		}
		Q.sorters_manager->NotifyBQPush();
	}
	Q.bq->mark_completed();
	Q.sorters_manager->NotifyQueueCompleted();

	for (auto& t : sorter_threads) t.join();
	completer.join();
	double t_end = now_s();
	if (times) { times[0] = t_end - t_begin; times[1] = t_sort_acc; }
	arena_cache = std::move(Q.memory_bins);
	return rc.load();
}

template <unsigned SIZE>
int sort_only(void* recs, void* tmp, uint64_t n, uint32_t key_bytes, int n_threads, int sort_kind, double* seconds)
{
	int64_t part = (256 * GetBufferWidth(sizeof(CKmer<SIZE>) / 8) + ALIGNMENT) * sizeof(CKmer<SIZE>);
	CMemoryPool pool(part * n_threads * MAGIC_NUMBER, part);
	double t0 = now_s();
	if (sort_kind == 0)
		RadulsSort::RadixSortMSD_AVX2<CKmer<SIZE>>((CKmer<SIZE>*)recs, (CKmer<SIZE>*)tmp, n, key_bytes - 1, n_threads, &pool);
	else {
		CSmallSort<SIZE>::Adjust(384);
		RadixSort::RadixSortMSD<CKmer<SIZE>, SIZE>((CKmer<SIZE>*)recs, (CKmer<SIZE>*)tmp, n, key_bytes - 1, n_threads, &pool);
	}
	if (seconds) *seconds = now_s() - t0;
	return (key_bytes % 2) ? 1 : 0;           // 1: result in tmp, 0: result in recs  (kb_sorter.h:776-779)
}

} // namespace

extern "C" {

// Runs n_bins bins through the reference's own stage 2 with n_sorters sorter threads.
// sort_kind: 0 = RADULS AVX2 (what the reference picks on Intel, kmc.h:1535-1551), 1 = radix.h + CSmallSort (non-Intel, kmc.h:1556-1560)
// times[0] = wall seconds for all bins, times[1] = seconds inside sort_func (summed over sorters)
int kmcref_process_bins(int k, int both_strands, uint32_t cutoff_min, uint32_t cutoff_max, uint32_t counter_max,
	uint32_t lut_prefix_len, int n_sorters, int sort_kind, int n_bins,
	const uint8_t* const* data, const uint64_t* size, const uint64_t* n_rec, const uint64_t* n_plus_x_recs,
	const uint64_t* const* pack_bytes, const uint64_t* const* pack_recs, const uint32_t* n_packs,
	uint8_t* const* out, const uint64_t* out_cap, uint64_t* out_bytes, uint64_t* const* lut, uint64_t* stats /* 4*n_bins */,
	double* times)
{
	std::vector<BinIO> bins(n_bins);
	for (int b = 0; b < n_bins; ++b) {
		bins[b] = BinIO{ data[b], size[b], n_rec[b], n_plus_x_recs[b], pack_bytes[b], pack_recs[b], n_packs[b],
			out[b], out_cap[b], 0, lut ? lut[b] : nullptr, {0, 0, 0, 0} };
	}
	int rc = -1;
	try {
		unsigned SIZE = (k + 31) / 32;
		switch (SIZE) {
		case 1: rc = run_bins<1>(k, both_strands, cutoff_min, cutoff_max, counter_max, lut_prefix_len, n_sorters, sort_kind, bins, times); break;
		case 2: rc = run_bins<2>(k, both_strands, cutoff_min, cutoff_max, counter_max, lut_prefix_len, n_sorters, sort_kind, bins, times); break;
		case 3: rc = run_bins<3>(k, both_strands, cutoff_min, cutoff_max, counter_max, lut_prefix_len, n_sorters, sort_kind, bins, times); break;
		case 4: rc = run_bins<4>(k, both_strands, cutoff_min, cutoff_max, counter_max, lut_prefix_len, n_sorters, sort_kind, bins, times); break;
		default: return -3;
		}
	} catch (const std::exception& e) {
		fprintf(stderr, "kmcref: %s\n", e.what());
		return -1;
	}
	for (int b = 0; b < n_bins; ++b) {
		out_bytes[b] = bins[b].out_bytes;
		for (int j = 0; j < 4; ++j) stats[4 * b + j] = bins[b].stats[j];
	}
	return rc;
}

// sort_func alone on n records of rec_words*8 bytes; returns 1 if the sorted data is in tmp, 0 if in recs, <0 on error
int kmcref_sort(void* recs, void* tmp, uint64_t n, uint32_t rec_words, uint32_t key_bytes, int n_threads, int sort_kind, double* seconds)
{
	try {
		switch (rec_words) {
		case 1: return sort_only<1>(recs, tmp, n, key_bytes, n_threads, sort_kind, seconds);
		case 2: return sort_only<2>(recs, tmp, n, key_bytes, n_threads, sort_kind, seconds);
		case 3: return sort_only<3>(recs, tmp, n, key_bytes, n_threads, sort_kind, seconds);
		case 4: return sort_only<4>(recs, tmp, n, key_bytes, n_threads, sort_kind, seconds);
		default: return -3;
		}
	} catch (const std::exception& e) {
		fprintf(stderr, "kmcref: %s\n", e.what());
		return -1;
	}
}

int kmcref_max_x(int k) { return (k % 32 == 0) ? 0 : MIN(31 - (k % 32), KMER_X); }

} // extern "C"
