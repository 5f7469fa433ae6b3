#!/usr/bin/env python
"""TEST INFRASTRUCTURE - applies the INTEGRATION.md patch to a COPY of the reference's kmc_core/kmc.h.

usage: patch_kmc.py <reference root> <scratch dir>
Writes <scratch>/kmc.h (patched) and <scratch>/kmc_runner.cpp (verbatim copy: the one translation unit that includes kmc.h, so that
its `#include "kmc.h"` finds the patched header first).  Nothing is written into the repository or the reference tree; the patch
is the ~15 lines below, guarded by KMC_WITH_B200 - without the define the copy compiles to the unmodified reference.

  1. kmc.h top:           #include "kb_sorter_b200.h"
  2. kmc.h:1564 (before CKmerQueue is created - it counts the sorter objects as writers): Params.n_sorters = number of GPU sorter
     objects = |KMC_B200_DEVICES| x KMC_B200_SORTERS_PER_GPU
  3. kmc.h:1576-1584:     construct CWKmerBinSorterB200<SIZE>(Params, Queues, device) instead of CWKmerBinSorter<SIZE>(Params, Queues, sort_func)
"""
import os
import shutil
import sys

ref, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)
src = open(os.path.join(ref, "kmc_core", "kmc.h")).read()


def replace_once(text, old, new):
    assert text.count(old) == 1, "patch anchor not found exactly once: %r" % old[:60]
    return text.replace(old, new)


src = replace_once(src, '#include "exception_aware_thread.h"\n',
                   '#include "exception_aware_thread.h"\n#ifdef KMC_WITH_B200\n#include "kb_sorter_b200.h"          // kmc_b200/host\n#endif\n')

src = replace_once(src, "\tQueues.kq = std::make_unique<CKmerQueue>(Params.n_bins, Params.n_sorters);\n",
                   """#ifdef KMC_WITH_B200
	// one or more sorter objects per GPU; they pull bins from the same CBinQueue (largest first) like the CPU sorters do
	std::vector<int> b200_devices = kmcb200_devices_from_env();      // KMC_B200_DEVICES="0,1,...", KMC_B200_SORTERS_PER_GPU (kb_sorter_b200.h)
	Params.n_sorters = (int)b200_devices.size();
#endif
	Queues.kq = std::make_unique<CKmerQueue>(Params.n_bins, Params.n_sorters);
""")

src = replace_once(src, """	vector<std::unique_ptr<CWKmerBinSorter<SIZE>>> w_sorters(Params.n_sorters);

	std::vector<CExceptionAwareThread> sorters_threads;

	for (int i = 0; i < Params.n_sorters; ++i)
	{
		w_sorters[i] = std::make_unique<CWKmerBinSorter<SIZE>>(Params, Queues, sort_func);
		sorters_threads.emplace_back(std::ref(*w_sorters[i].get()));
	}
""", """	std::vector<CExceptionAwareThread> sorters_threads;
#ifdef KMC_WITH_B200
	vector<std::unique_ptr<CWKmerBinSorterB200<SIZE>>> w_sorters(Params.n_sorters);
	for (int i = 0; i < Params.n_sorters; ++i)
	{
		w_sorters[i] = std::make_unique<CWKmerBinSorterB200<SIZE>>(Params, Queues, b200_devices[i]);
		sorters_threads.emplace_back(std::ref(*w_sorters[i].get()));
	}
#else
	vector<std::unique_ptr<CWKmerBinSorter<SIZE>>> w_sorters(Params.n_sorters);
	for (int i = 0; i < Params.n_sorters; ++i)
	{
		w_sorters[i] = std::make_unique<CWKmerBinSorter<SIZE>>(Params, Queues, sort_func);
		sorters_threads.emplace_back(std::ref(*w_sorters[i].get()));
	}
#endif
""")
open(os.path.join(out, "kmc.h"), "w").write(src)
shutil.copyfile(os.path.join(ref, "kmc_core", "kmc_runner.cpp"), os.path.join(out, "kmc_runner.cpp"))
print("patched kmc.h ->", os.path.join(out, "kmc.h"))
