// kmc_b200 — host-side drop-in for KMC's per-bin stage 2 (C++14, header only, to be compiled INSIDE a KMC tree).
//
// CKmerBinSorterB200<SIZE> has the constructor and ProcessBins() shape of the reference's CKmerBinSorter<SIZE>
// (kmc_core/kb_sorter.h:42-237) and speaks to the same queues, so CKMC<SIZE>::ProcessStage2_impl only has to
// construct it instead of CWKmerBinSorter where the sorter threads are created (kmc_core/kmc.h:1576-1584);
// KMC::Runner, Stage1Params/Stage2Params and every other class stay untouched.
//
// What it does per bin (everything between sorters_manager->GetNext and kq->push, kb_sorter.h:216-233, :1273):
//   inputs   bin bytes in the CMemoryBins arena (mba_input_file), n_rec from CBinQueue, expander packs from
//            CExpanderPackDesc (queues.h:376-396)
//   GPU      kmcb200_process_bin: H2D, expand, radix sort, count/compact, D2H       (include/kmc_b200.h)
//   outputs  out_buffer (mba_suffix), one data pack (0, out_bytes), raw LUT (mba_lut), the four counters -> kq->push
// Errors follow the reference's convention: CCriticalErrorHandler::Inst().HandleCriticalError(msg)
// (critical_error_handler.h:73-79) - there is no CPU fallback.
//
// The bin bytes are copied straight out of the arena; registering the arena once with
// cudaHostRegister(buffer, total_size) makes those copies asynchronous DMA ("pinned arena").
#ifndef KMC_B200_KB_SORTER_B200_H
#define KMC_B200_KB_SORTER_B200_H

#include "defs.h"
#include "params.h"
#include "kmer.h"
#include "critical_error_handler.h"
#include "kmc_b200.h"

#include <algorithm>
#include <cstdlib>
#include <list>
#include <sstream>
#include <string>
#include <vector>

// Which GPU each sorter object of this run drives: KMC_B200_DEVICES="0,1,2,..." (default "0"), each repeated
// KMC_B200_SORTERS_PER_GPU times (default 1; 2 overlaps one bin's PCIe copies with the other's kernels).  One entry = one
// CKmerBinSorterB200 object = one writer of CKmerQueue (kmc.h:1564), all pulling from the same CBinQueue in get_sorted_req_sizes order.
inline std::vector<int> kmcb200_devices_from_env()
{
	std::vector<int> devs;
	const char* e = std::getenv("KMC_B200_DEVICES");
	std::string s = e && *e ? e : "0";
	size_t pos = 0;
	while (pos < s.size())
	{
		size_t c = s.find(',', pos);
		if (c == std::string::npos) c = s.size();
		if (c > pos) devs.push_back(std::atoi(s.substr(pos, c - pos).c_str()));
		pos = c + 1;
	}
	if (devs.empty()) devs.push_back(0);
	int per = 1;
	if (const char* p = std::getenv("KMC_B200_SORTERS_PER_GPU")) per = std::max(1, std::min(4, std::atoi(p)));
	std::vector<int> out;
	for (int r = 0; r < per; ++r)
		for (int d : devs) out.push_back(d);
	return out;
}

template <unsigned SIZE> class CKmerBinSorterB200
{
	CBinDesc* bd;
	CExpanderPackDesc* epd;
	CKmerQueue* kq;
	CMemoryBins* memory_bins;
	CSortersManager* sorters_manager;
	uint32 max_x;
	uint32 lut_prefix_len;
	kmcb200_ctx* ctx;
	uint64 sum_n_rec = 0, sum_n_plus_x_rec = 0;

	void fail(const char* what)
	{
		std::ostringstream ostr;
		ostr << "Error: kmc_b200 " << what << ": " << kmcb200_last_error(ctx);
		CCriticalErrorHandler::Inst().HandleCriticalError(ostr.str());
	}

public:
	// same first two arguments as CKmerBinSorter's constructor (kb_sorter.h:165); the sort_func argument is gone
	CKmerBinSorterB200(CKMCParams& Params, CKMCQueues& Queues, int device = 0) : ctx(nullptr)
	{
		bd = Queues.bd.get();
		epd = Queues.epd.get();
		kq = Queues.kq.get();
		memory_bins = Queues.memory_bins.get();
		sorters_manager = Queues.sorters_manager.get();
		max_x = Params.max_x;
		lut_prefix_len = Params.lut_prefix_len;

		if (Params.output_type != OutputType::KMC || Params.without_output)
			CCriticalErrorHandler::Inst().HandleCriticalError("Error: kmc_b200 supports the KMC database output only");
		kmcb200_params p;
		p.kmer_len = (uint32_t)Params.kmer_len;
		p.both_strands = Params.both_strands ? 1u : 0u;
		p.cutoff_min = (uint32_t)Params.cutoff_min;
		p.cutoff_max = (uint32_t)Params.cutoff_max;          // kb_sorter.h:186
		p.counter_max = (uint32_t)Params.counter_max;        // kb_sorter.h:187
		p.lut_prefix_len = Params.lut_prefix_len;
		p.device = device;
		p.n_slots = 1;
		if (kmcb200_create(&p, &ctx) != KMCB200_OK)
		{
			std::ostringstream ostr;
			ostr << "Error: kmc_b200 cannot start on device " << device << ": " << kmcb200_last_error(nullptr);
			CCriticalErrorHandler::Inst().HandleCriticalError(ostr.str());
		}
	}

	~CKmerBinSorterB200() { kmcb200_destroy(ctx); }

	// kb_sorter.h:118-122 (read by CKMC::ProcessStage2_impl for its statistics, kmc.h:1736-1741)
	void GetDebugStats(uint64& _sum_n_recs, uint64& _sum_n_plus_x_recs)
	{
		_sum_n_recs = sum_n_rec;
		_sum_n_plus_x_recs = sum_n_plus_x_rec;
	}

	CKmerBinSorterB200(const CKmerBinSorterB200&) = delete;
	CKmerBinSorterB200& operator=(const CKmerBinSorterB200&) = delete;

	// kb_sorter.h:210-237
	void ProcessBins()
	{
		int32 bin_id;
		uchar* data;
		uint64 size, n_rec;
		int n_sorting_threads;
		std::vector<uint64_t> pack_bytes;

		while (sorters_manager->GetNext(bin_id, data, size, n_rec, n_sorting_threads))
		{
			CMemDiskFile* file;
			std::string desc;
			uint64 tmp_size, tmp_n_rec, n_plus_x_recs;
			bd->read(bin_id, file, desc, tmp_size, tmp_n_rec, n_plus_x_recs);
			sum_n_rec += n_rec;
			sum_n_plus_x_rec += n_plus_x_recs;

			std::list<std::pair<uint64, uint64>> packs;
			epd->pop(bin_id, packs);
			pack_bytes.clear();
			for (auto& p : packs)
				pack_bytes.push_back(p.first);

			uchar* out_buffer = nullptr;
			uchar* raw_lut = nullptr;
			memory_bins->reserve(bin_id, out_buffer, CMemoryBins::mba_suffix);
			memory_bins->reserve(bin_id, raw_lut, CMemoryBins::mba_lut);

			const uint64 lut_size = (1ull << (2 * lut_prefix_len)) * sizeof(uint64);
			const uint64 out_capacity = kmcb200_out_capacity(ctx, n_rec);        // what kb_reader.h:141-150 reserved
			uint64_t out_bytes = 0;
			uint64_t stats[4] = { 0, 0, 0, 0 };

			// mba_suffix/mba_lut may overlay the input file inside the arena (queues.h:468-484): kmcb200_process_bin
			// has copied the whole bin to the GPU before it writes a single output byte, so the overlay is harmless.
			int rc = kmcb200_process_bin(ctx, bin_id, data, size, n_rec, n_plus_x_recs,
				pack_bytes.empty() ? nullptr : pack_bytes.data(), nullptr, (uint32_t)pack_bytes.size(),
				out_buffer, out_capacity, &out_bytes, (uint64_t*)raw_lut, stats);
			if (rc != KMCB200_OK)
				fail("stage 2 failed");

			memory_bins->free(bin_id, CMemoryBins::mba_input_file);               // kb_sorter.h:225 (also for empty bins)

			std::list<std::pair<uint64, uint64>> data_packs;
			data_packs.emplace_back(0, out_bytes);                                // kb_sorter.h:1269-1271
			kq->push(bin_id, out_buffer, data_packs, raw_lut, lut_size, stats[0], stats[1], stats[2], stats[3]);   // :1273

			memory_bins->free(bin_id, CMemoryBins::mba_input_array);              // :1275-1279 (never touched by the GPU path)
			memory_bins->free(bin_id, CMemoryBins::mba_tmp_array);
			if (max_x)
				memory_bins->free(bin_id, CMemoryBins::mba_kxmer_counters);       // :1109

			sorters_manager->ReturnThreads(n_sorting_threads, bin_id);            // :233
		}
		kq->mark_completed();                                                      // :236
	}
};

// wrapper with the shape of CWKmerBinSorter (kb_sorter.h:1298-1322) so that it can be handed to a thread
template <unsigned SIZE> class CWKmerBinSorterB200
{
	std::unique_ptr<CKmerBinSorterB200<SIZE>> kbs;

public:
	CWKmerBinSorterB200(CKMCParams& Params, CKMCQueues& Queues, int device = 0)
	{
		kbs = std::make_unique<CKmerBinSorterB200<SIZE>>(Params, Queues, device);
	}
	void GetDebugStats(uint64& _sum_n_recs, uint64& _sum_n_plus_x_recs) { kbs->GetDebugStats(_sum_n_recs, _sum_n_plus_x_recs); }
	void operator()() { kbs->ProcessBins(); }
};

#endif
