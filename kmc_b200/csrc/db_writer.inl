// kmc_b200 — host side of SURVEY section 8f N3: assembling the KMC database files from per-bin GPU results without the reference's
// single-threaded completer loop (kmc_core/kb_completer.cpp:59-326: fwrite of every bin's records, a scalar prefix sum over its LUT,
// the signature map, the footer).  Included by kmc_b200.cu.
//
//   * the LUT prefix sum with the running record count (kb_completer.cpp:191-201) is done on the GPU (lut_scan_kernel, behind
//     kmcb200_wait_bin_scanned): the host receives the 4^p uint64 exactly as they go into .kmc_pre;
//   * the emitted records are copied device -> host straight into a PINNED STAGING RING owned by the writer (kmcb200_db_reserve), and a
//     writer thread appends committed regions to .kmc_suf / .kmc_pre in commit order while the GPU works on the next bins;
//   * kmcb200_db_close writes what ProcessBinsSecondStage writes (kb_completer.cpp:284-320): total records, signature map, header, markers.
// The files are byte-identical to the reference's for the same bins in the same order (tests/test_db_writer.py replays a database
// written by the reference CLI through this writer and compares the bytes).
#include <condition_variable>
#include <deque>
#include <mutex>

struct kmcb200_db_writer {
	kmcb200_db_params prm{};
	FILE* f_pre = nullptr;
	FILE* f_suf = nullptr;
	std::string err;
	// pinned staging ring: regions are reserved by the producer (one at a time, in commit order) and released by the writer thread
	uint8_t* ring = nullptr;
	uint64_t ring_bytes = 0, head = 0, tail = 0, used = 0;          // [tail, head) is in use (modulo ring_bytes); a region never wraps
	struct Region { uint8_t* ptr; uint64_t bytes, ring_advance; bool own, pinned; };
	Region open_region{nullptr, 0, 0, false, false};
	bool ring_pinned = false;
	bool region_open = false;
	struct Job { Region payload; uint64_t payload_bytes; std::vector<uint64_t> lut; };
	std::deque<Job> jobs;
	std::mutex mtx;
	std::condition_variable cv_jobs, cv_space;
	std::thread thread;
	bool closing = false, io_failed = false;
	// running totals (kb_completer.cpp:206-209) and the signature map (:211-221)
	uint64_t n_recs = 0, n_unique = 0, n_cutoff_min = 0, n_cutoff_max = 0, n_total = 0;
	uint32_t lut_pos = 0;
	std::vector<uint32_t> sig_map;
};

namespace {

// pinned when a CUDA device is present (the D2H target), plain memory otherwise (host-only use of the writer: tests/test_db_writer.py)
uint8_t* db_alloc(uint64_t bytes, bool* pinned)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) == cudaSuccess) { *pinned = true; return static_cast<uint8_t*>(p); }
	cudaGetLastError();
	*pinned = false;
	return static_cast<uint8_t*>(malloc(bytes));
}
void db_free(uint8_t* p, bool pinned) { if (!p) return; if (pinned) cudaFreeHost(p); else free(p); }

void db_store_uint(FILE* f, uint64_t x, int bytes) { for (int i = 0; i < bytes; ++i) fputc((int)((x >> (8 * i)) & 0xFF), f); }      // little endian, as kb_completer's store_uint

void db_writer_loop(kmcb200_db_writer* w)
{
	for (;;) {
		kmcb200_db_writer::Job job;
		{
			std::unique_lock<std::mutex> lk(w->mtx);
			w->cv_jobs.wait(lk, [&] { return !w->jobs.empty() || w->closing; });
			if (w->jobs.empty()) return;
			job = std::move(w->jobs.front());
			w->jobs.pop_front();
		}
		bool ok = true;
		if (job.payload_bytes) ok = fwrite(job.payload.ptr, 1, job.payload_bytes, w->f_suf) == job.payload_bytes;          // kb_completer.cpp:154-170
		if (ok && !job.lut.empty()) ok = fwrite(job.lut.data(), sizeof(uint64_t), job.lut.size(), w->f_pre) == job.lut.size();   // :200
		{
			std::lock_guard<std::mutex> lk(w->mtx);
			if (!ok) w->io_failed = true;
			if (job.payload.own) db_free(job.payload.ptr, job.payload.pinned);
			else { w->tail = (w->tail + job.payload.ring_advance) % w->ring_bytes; w->used -= job.payload.ring_advance; }
		}
		w->cv_space.notify_all();
	}
}

__global__ void __launch_bounds__(1024) lut_scan_kernel(uint64_t* lut, uint64_t n, uint64_t base)
{
	// exclusive prefix sum in place, offset by `base` (kb_completer.cpp:191-201); one CTA, 4^p <= 2^30 entries in chunks of 1024
	__shared__ uint64_t s_w[32];
	__shared__ uint64_t carry;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) carry = base;
	__syncthreads();
	for (uint64_t i0 = 0; i0 < n; i0 += 1024) {
		const uint64_t i = i0 + tid;
		const uint64_t v = i < n ? lut[i] : 0;
		uint64_t inc = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) { const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += t; }
		if (lane == 31) s_w[warp] = inc;
		__syncthreads();
		uint64_t b = carry;
		for (uint32_t w = 0; w < warp; ++w) b += s_w[w];
		if (i < n) lut[i] = b + inc - v;
		__syncthreads();
		if (tid == 1023) carry = b + inc;
		__syncthreads();
	}
}

int db_fail(kmcb200_db_writer* w, int code, const char* fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	if (w) w->err = buf;
	else g_create_error = buf;
	return code;
}

}  // namespace

extern "C" {

int kmcb200_db_open(const kmcb200_db_params* prm, const char* path_prefix, uint64_t staging_bytes, kmcb200_db_writer** out)
{
	if (!prm || !path_prefix || !out) return db_fail(nullptr, KMCB200_ERR_INVALID, "null argument");
	*out = nullptr;
	if (prm->signature_len < 5 || prm->signature_len > 11 || prm->lut_prefix_len < 1 || prm->lut_prefix_len > 15)
		return db_fail(nullptr, KMCB200_ERR_INVALID, "signature_len %u / lut_prefix_len %u out of range", prm->signature_len, prm->lut_prefix_len);
	kmcb200_db_writer* w = new kmcb200_db_writer();
	w->prm = *prm;
	const std::string base(path_prefix);
	w->f_pre = fopen((base + ".kmc_pre").c_str(), "wb");
	w->f_suf = fopen((base + ".kmc_suf").c_str(), "wb");
	if (!w->f_pre || !w->f_suf) {
		if (w->f_pre) fclose(w->f_pre);
		if (w->f_suf) fclose(w->f_suf);
		delete w;
		return db_fail(nullptr, KMCB200_ERR_INVALID, "cannot create %s.kmc_pre / .kmc_suf", path_prefix);
	}
	setvbuf(w->f_suf, nullptr, _IOFBF, 1 << 24);
	fwrite("KMCP", 1, 4, w->f_pre);          // markers at the beginning (kb_completer.cpp:121-127)
	fwrite("KMCS", 1, 4, w->f_suf);
	w->ring_bytes = std::max<uint64_t>(staging_bytes, 1 << 20);
	w->ring = db_alloc(w->ring_bytes, &w->ring_pinned);
	if (!w->ring) {
		fclose(w->f_pre); fclose(w->f_suf);
		delete w;
		return db_fail(nullptr, KMCB200_ERR_CUDA, "cannot allocate %llu bytes of pinned staging memory", (unsigned long long)staging_bytes);
	}
	w->sig_map.assign(((size_t)1 << (2 * prm->signature_len)) + 1, 0u);
	w->thread = std::thread(db_writer_loop, w);
	*out = w;
	return 0;
}

const char* kmcb200_db_last_error(const kmcb200_db_writer* w) { return w ? w->err.c_str() : g_create_error.c_str(); }
uint64_t kmcb200_db_records(const kmcb200_db_writer* w) { return w ? w->n_recs : 0; }

// A pinned region for the next bin's emitted records (the D2H target: pass it as out_suffix to kmcb200_submit_bin).  One region is open at
// a time; it is committed (or dropped with bytes = 0) by kmcb200_db_commit_bin.  Blocks while the ring is full.
int kmcb200_db_reserve(kmcb200_db_writer* w, uint64_t bytes, uint8_t** out_ptr)
{
	if (!w || !out_ptr) return KMCB200_ERR_INVALID;
	if (w->region_open) return db_fail(w, KMCB200_ERR_BUSY, "a region is already open");
	const uint64_t need = (std::max<uint64_t>(bytes, 1) + 255) & ~255ull;
	if (need > w->ring_bytes / 2) {          // larger than half of the ring: its own pinned block, freed after it is written
		bool pinned = false;
		uint8_t* p = db_alloc(need, &pinned);
		if (!p) return db_fail(w, KMCB200_ERR_CUDA, "cannot allocate %llu staging bytes", (unsigned long long)need);
		w->open_region = {p, need, 0, true, pinned};
	} else {
		std::unique_lock<std::mutex> lk(w->mtx);
		uint64_t advance = 0, start = 0;
		w->cv_space.wait(lk, [&] {
			start = w->head; advance = need;
			if (start + need > w->ring_bytes) { advance = (w->ring_bytes - start) + need; start = 0; }          // does not fit behind the head: skip to the start
			return w->used + advance <= w->ring_bytes;
		});
		w->open_region = {w->ring + start, need, advance, false, false};
		w->head = (w->head + advance) % w->ring_bytes;
		w->used += advance;
	}
	w->region_open = true;
	*out_ptr = w->open_region.ptr;
	return 0;
}

// Hands the open region (its first payload_bytes bytes) and the bin's LUT to the writer thread.  scanned_lut: 4^p entries, already the
// exclusive prefix sum offset by kmcb200_db_records() (kmcb200_wait_bin_scanned delivers exactly that); raw_lut != 0: the entries are raw
// counts and the prefix sum is done here on the host (a caller without a GPU-side scan).  signatures: the minimizer signatures that stage
// 1 mapped to this bin (CSignatureMapper) - they get the bin's ordinal in the file (kb_completer.cpp:211-221).
int kmcb200_db_commit_bin(kmcb200_db_writer* w, uint64_t payload_bytes, const uint64_t* lut, int raw_lut, const uint64_t stats[4],
	const uint32_t* signatures, uint32_t n_signatures)
{
	if (!w || !lut || !stats) return KMCB200_ERR_INVALID;
	if (!w->region_open) return db_fail(w, KMCB200_ERR_INVALID, "no open region");
	if (payload_bytes > w->open_region.bytes) return db_fail(w, KMCB200_ERR_CAPACITY, "payload larger than the reserved region");
	const uint32_t rec = (w->prm.kmer_len - w->prm.lut_prefix_len) / 4 + w->prm.counter_size;
	if (rec && payload_bytes % rec) return db_fail(w, KMCB200_ERR_INVALID, "payload is not a whole number of %u-byte records", rec);
	const uint64_t n_lut = 1ull << (2 * w->prm.lut_prefix_len);
	kmcb200_db_writer::Job job;
	job.payload = w->open_region;
	job.payload_bytes = payload_bytes;
	job.lut.assign(lut, lut + n_lut);
	const uint64_t bin_recs = rec ? payload_bytes / rec : 0;
	if (raw_lut) { uint64_t acc = w->n_recs; for (uint64_t i = 0; i < n_lut; ++i) { const uint64_t x = job.lut[i]; job.lut[i] = acc; acc += x; } }
	else if (n_lut && job.lut[0] != w->n_recs) return db_fail(w, KMCB200_ERR_INVALID, "scanned LUT starts at %llu but %llu records precede this bin", (unsigned long long)job.lut[0], (unsigned long long)w->n_recs);
	w->n_recs += bin_recs;
	w->n_unique += stats[0]; w->n_cutoff_min += stats[1]; w->n_cutoff_max += stats[2]; w->n_total += stats[3];
	for (uint32_t i = 0; i < n_signatures; ++i) if (signatures[i] < w->sig_map.size()) w->sig_map[signatures[i]] = w->lut_pos;
	++w->lut_pos;
	w->region_open = false;
	{
		std::lock_guard<std::mutex> lk(w->mtx);
		if (w->io_failed) return db_fail(w, KMCB200_ERR_INVALID, "writing the database files failed");
		w->jobs.push_back(std::move(job));
	}
	w->cv_jobs.notify_one();
	return 0;
}

// Drains the writer thread and writes the footer (kb_completer.cpp:284-320).  totals (optional): n_unique, n_cutoff_min, n_cutoff_max, n_total.
int kmcb200_db_close(kmcb200_db_writer* w, uint64_t totals[4])
{
	if (!w) return KMCB200_ERR_INVALID;
	{
		std::lock_guard<std::mutex> lk(w->mtx);
		w->closing = true;
	}
	w->cv_jobs.notify_all();
	if (w->thread.joinable()) w->thread.join();
	if (w->region_open && w->open_region.own) db_free(w->open_region.ptr, w->open_region.pinned);
	int rc = w->io_failed ? KMCB200_ERR_INVALID : 0;
	fwrite("KMCS", 1, 4, w->f_suf);                                  // marker at the end
	if (fclose(w->f_suf) != 0) rc = KMCB200_ERR_INVALID;
	FILE* f = w->f_pre;
	fwrite(&w->n_recs, 1, sizeof(uint64_t), f);
	fwrite(w->sig_map.data(), sizeof(uint32_t), w->sig_map.size(), f);
	uint32_t offset = 0;
	db_store_uint(f, w->prm.kmer_len, 4); offset += 4;
	db_store_uint(f, 0, 4); offset += 4;                             // mode 0 (counting)
	db_store_uint(f, w->prm.counter_size, 4); offset += 4;
	db_store_uint(f, w->prm.lut_prefix_len, 4); offset += 4;
	db_store_uint(f, w->prm.signature_len, 4); offset += 4;
	db_store_uint(f, w->prm.cutoff_min, 4); offset += 4;
	db_store_uint(f, w->prm.cutoff_max, 4); offset += 4;
	db_store_uint(f, w->n_unique - w->n_cutoff_min - w->n_cutoff_max, 8); offset += 8;
	db_store_uint(f, w->prm.both_strands ? 0 : 1, 1); offset += 1;
	for (int i = 0; i < 27; ++i) { db_store_uint(f, 0, 1); offset += 1; }
	db_store_uint(f, 0x200, 4); offset += 4;
	db_store_uint(f, offset, 4);
	fwrite("KMCP", 1, 4, f);
	if (fclose(f) != 0) rc = KMCB200_ERR_INVALID;
	if (totals) { totals[0] = w->n_unique; totals[1] = w->n_cutoff_min; totals[2] = w->n_cutoff_max; totals[3] = w->n_total; }
	db_free(w->ring, w->ring_pinned);
	delete w;
	return rc;
}

// kmcb200_wait_bin, but the LUT arrives as the completer writes it: exclusive prefix sum of the bin's raw counts offset by lut_base (the
// records that precede the bin in the file), computed on the GPU right before the copy.
int kmcb200_wait_bin_scanned(kmcb200_ctx* ctx, uint32_t slot, uint64_t lut_base, uint64_t* out_bytes, uint64_t stats[4])
{
	if (int rc = check_slot(ctx, slot)) return rc;
	Slot& s = ctx->slots[slot];
	if (!s.busy) return fail(ctx, KMCB200_ERR_INVALID, "slot %u has no submitted bin", slot);
	s.scan_lut = true; s.scan_base = lut_base;
	return kmcb200_wait_bin(ctx, slot, out_bytes, stats);
}

}  // extern "C"
