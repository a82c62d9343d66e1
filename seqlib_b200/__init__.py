"""seqlib_b200 -- host-side bindings of the B200-native seed-and-extend / local-assembly engine (include/seqlib_b200.h).

The product is the CUDA library seqlib_b200/libseqlib_b200.so (sources in csrc/) and the C++ drop-in classes in
libSeqLibB200.so (cxx/, headers in include/SeqLib/).  The Python modules only bind the C ABI for tests and benchmarks:

    capi   index construction / load / write, b200_mem_align_batch, ksw batch, the fermi-lite half (count, correct, assemble,
           assemble_windows, FMD-index)
    fastq  FASTA/FASTQ ingest: stream parser (kseq-exact) and the device parser for strict four-line FASTQ
    sam    SAM text of an alignment batch (bwa's mem_reg2sam)
    synth  synthetic references / reads of the benchmark configurations
    shard  read sharding for one-process-per-GPU runs

There is no CPU fallback: every compute call fails with B200_ERR_CUDA when no sm_100 device is usable.
"""
