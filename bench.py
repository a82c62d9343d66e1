#!/usr/bin/env python
"""bench.py -- reads/s of the seed-and-extend hot path (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's own bwa C on the host cores)

A "step" is one pass of the hot path (SMEM seeding -> chaining -> banded SW -> CIGAR/MAPQ, i.e. what
BWAAligner::alignSequence computes) over one batch of synthetic 150-bp reads per GPU.  Default workload is
BASELINE.json configs[1]: 10 M reads vs a 3 Gb uniform-random reference (24 contigs) on one B200; with
N > 1 every GPU gets its own 10 M reads (weak scaling), the index is built once on rank 0 and broadcast
over NCCL, the read shards are scattered over NCCL.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "ksw"])
    ap.add_argument("--ref-len", type=int, default=int(os.environ.get("B200_BENCH_REF_LEN", 3_000_000_000)))
    ap.add_argument("--reads", type=int, default=int(os.environ.get("B200_BENCH_READS", 10_000_000)), help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary measurements (config 4: FermiAssembler::PerformAssembly)")
    ap.add_argument("--parity-reads", type=int, default=int(os.environ.get("B200_BENCH_PARITY_READS", 500_000)),
                    help="reads of the timed batch re-aligned by the reference library and compared bit for bit (N=1)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 7:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def chunk_plan(n, ch):
    """The chunk sizes b200_batch_run uses (csrc/engine.cu): a quarter-size first chunk, full chunks, a tapering tail."""
    taper = os.environ.get("B200_CHUNK_TAPER", "0")
    if taper == "2" and n > 3 * ch:
        first, last = (ch * 3 // 4) & ~1023, (ch // 2) & ~1023
        mid_total = n - first - last
        k = max(1, (mid_total * 4 + ch * 5 - 1) // (ch * 5))
        mid = ((mid_total + k - 1) // k + 1023) & ~1023
        plan, rem = [first], n - first
        while rem > last:
            t = min(mid, rem - last); plan.append(t); rem -= t
        plan.append(rem)
        return plan
    if taper in ("0", "2"):
        return [min(ch, n - i) for i in range(0, n, ch)]
    plan, rem, q = [], n, max(ch // 4, 1024)
    if rem > ch:
        plan.append(q); rem -= q
    while rem > ch + ch // 2:
        plan.append(ch); rem -= ch
    while rem > q:
        t = min(rem, max(q, (rem // 2 + 1023) & ~1023)); plan.append(t); rem -= t
    if rem > 0:
        plan.append(rem)
    return plan


def seed_traffic_per_read():
    """DRAM bytes per read of the seeding kernel from the committed ncu --set full capture (None if absent)."""
    p = os.path.join(ROOT, "profiles", "r02_seed_traffic.json")
    try:
        d = json.load(open(p))
        return (d["dram_bytes_read"] + d["dram_bytes_write"]) / d["reads_in_launch"], d["source"]
    except Exception:
        return None, None


def fermi_extra(args, with_cpu):
    """Config 4 (FermiAssembler::PerformAssembly, 1M x 150bp at 150x): secondary numbers carried in the same JSON line."""
    from seqlib_b200 import capi, synth
    n = int(os.environ.get("B200_BENCH_ASM_READS", 1_000_000))
    pac = synth.reference(n, seed=0x5EED0005)
    ctg = synth.contigs_for(n, 1, "asm")
    seqs, off, _, _ = synth.reads(pac, n, ctg, n, 150, 0.01, 0.0, seed=0x5EED0006)
    quals = np.full(len(seqs), ord("I"), dtype=np.uint8)
    opt = capi.fml_default_opt()
    capi.fml_assemble_flat(opt, seqs, quals, off)          # warm-up (allocations, first launches)
    ts, st = [], None
    for _ in range(2):
        t0 = time.perf_counter()
        utgs = capi.fml_assemble_flat(opt, seqs, quals, off)
        ts.append(time.perf_counter() - t0)
        st = capi.fml_last_stats()
    out = {"workload": "fml_assemble: %d x 150bp reads from a %d bp random region (150x), 1%% substitutions" % (n, n),
           "e2e_reads_per_s": n / float(np.mean(ts)), "e2e_seconds": float(np.mean(ts)),
           "stage_ms": {k: st[k] for k in ("ms_count", "ms_ec", "ms_flt", "ms_fmd", "ms_nodes", "ms_walk_host", "ms_clean_host")},
           "n_utg": len(utgs), "longest_utg": max([len(u["seq"]) for u in utgs] + [0]), "gpu_launches": st["n_launches"],
           "ec_table_probes": st["n_lookups"], "fmd_symbols": st["fmd_symbols"]}
    if with_cpu:
        try:
            from oracle import pyref_fml
            m = 50_000
            pac2 = synth.reference(m, seed=0x5EED0005)
            s2, o2, _, _ = synth.reads(pac2, m, synth.contigs_for(m, 1, "asm"), m, 150, 0.01, 0.0, seed=0x5EED0006)
            exp, sec = pyref_fml.assemble(pyref_fml.default_opt(), s2, np.full(len(s2), ord("I"), dtype=np.uint8), o2)
            out["cpu_baseline"] = {"value": m / sec, "unit": "reads/s", "cores": 1, "kind": "reference",
                                   "sample": "fml_assemble (fermi-lite/misc.c:280-302), n_threads=1, %d reads at the same coverage, %.1f s" % (m, sec)}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "sample": "unavailable: %s" % e}
        try:
            # parity at the config's own size: the reference's fml_assemble on the SAME 1 M reads with all host threads (its unitig SET does
            # not depend on the thread count), compared as canonical (min of sequence / reverse complement) sorted sequences
            from oracle import pyref_fml
            ropt = pyref_fml.default_opt()
            ropt.n_threads = os.cpu_count() or 1
            exp, sec = pyref_fml.assemble(ropt, seqs, quals, off)
            comp = bytes.maketrans(b"ACGTacgt", b"TGCAtgca")

            def canon(us):
                return sorted(min(u["seq"], u["seq"].translate(comp)[::-1]) for u in us)
            a, b = canon(utgs), canon(exp)
            out["parity"] = {"n_reads": n, "unitigs": len(b), "mismatches": int(a != b) and (len(set(a) ^ set(b)) or 1),
                             "against": "oracle/_ref fml_assemble (fermi-lite/misc.c:280-302), %d threads, %.1f s" % (ropt.n_threads, sec)}
        except Exception as e:
            out["parity"] = {"error": str(e)}
    return out


def make_workload(args, n_total_reads):
    from seqlib_b200 import synth
    l_pac = args.ref_len
    n_ctg = 24 if l_pac >= 24 * 1000 else 1
    t0 = time.time()
    pac = synth.reference(l_pac)
    ctg = synth.contigs_for(l_pac, n_ctg)
    seqs, off, pos, strand = synth.reads(pac, l_pac, ctg, n_total_reads, args.read_len, 0.01, 0.0)
    make_workload.truth = (pos, strand)
    return pac, ctg, seqs, off, time.time() - t0


def parity_block(res, seqs, off, ids, ctg, ridx, opt, n_par, cores):
    """Bit-exactness inside the headline run (BASELINE.md parity gate): (a) every read's primary hit against the simulated
    origin, (b) the first n_par reads through the reference's own mem_align1 + mem_reg2aln (oracle/_ref, bwa/bwamem_extra.c:103-115,
    bwa/bwamem.c:1119-1189) on the SAME index arrays, compared field by field with the GPU hits (regions, CIGAR, MD, MAPQ)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity
    pos, strand = make_workload.truth
    n = len(off) - 1
    rec, mapped = parity.truth_recovery(res, pos[:n], strand[:n], ctg)
    out = {"truth_within_8bp": rec, "mapped_fraction": mapped}
    if ridx is not None:
        from oracle import pyref
        n_par = int(min(n, n_par))
        L = int(off[1] - off[0])
        exp, sec = pyref.align(ridx, (seqs[:n_par * L], off[:n_par + 1]), opt, ids[:n_par], n_threads=cores)
        bad, msgs = parity.compare_prefix(res, exp, n_par)
        out.update({"n": n_par, "mismatches": int(bad), "hits_compared": int(exp.hit_off[-1]), "reference_seconds": sec,
                    "against": "oracle/_ref mem_align1 + mem_reg2aln on the same index arrays, %d threads" % cores,
                    "messages": msgs[:3]})
    return out


def cpu_baseline_sample(ridx, seqs, off, read_len, opt, target_s, cores):
    """The reference's own batched schedule (mem_process_seqs, all cores) on a bounded sample of the same reads."""
    from oracle import pyref
    n_all = len(off) - 1
    probe = min(n_all, 4000)
    t = pyref.process_seqs(ridx, (seqs[:probe * read_len], off[:probe + 1]), opt, cores)
    rate = probe / max(t, 1e-6)
    n = int(min(n_all, max(probe, rate * target_s)))
    t = pyref.process_seqs(ridx, (seqs[:n * read_len], off[:n + 1]), opt, cores)
    return n / t, n, t


def run_reference_arm(args):
    """--impl reference: the reference's bwa C (oracle/_ref, compiled from the mount) on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyref
    from seqlib_b200 import capi
    cores = os.cpu_count() or 1
    capi.set_device(0)
    pac, ctg, seqs, off, t_gen = make_workload(args, min(args.reads, 2_000_000))
    # the 3 Gb FM-index is a pure function of the text; the GPU builder only prepares the input of the timed reference code
    t0 = time.time()
    idx = capi.Index.construct_pac(pac, args.ref_len, ctg, keep_host=True)
    t_index = time.time() - t0
    ridx = pyref.RefIndex.from_view(idx.view(), keep=idx)
    opt = pyref.default_opt()
    probe = min(len(off) - 1, 4000)
    t = pyref.process_seqs(ridx, (seqs[:probe * args.read_len], off[:probe + 1]), opt, cores)
    rate = probe / max(t, 1e-6)
    n = int(min(len(off) - 1, max(probe, rate * args.cpu_seconds)))
    sub = (seqs[:n * args.read_len], off[:n + 1])
    for _ in range(max(0, min(args.warmup, 1))):
        pyref.process_seqs(ridx, sub, opt, cores)
    times = [pyref.process_seqs(ridx, sub, opt, cores) for _ in range(args.steps)]
    tot = sum(times)
    value = n * args.steps / tot
    line = {
        "impl": "reference", "metric": "150bp reads/sec (seed+chain+SW end-to-end)", "value": value, "unit": "reads/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "10M x 150bp synthetic reads vs %d bp random reference (24 contigs)" % args.ref_len,
                   "reads_per_step": n, "read_len": args.read_len, "ref_len": args.ref_len, "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": "reference",
                         "sample": "%d reads per step through mem_process_seqs (bwa/bwamem.c:1235-1264), %d threads" % (n, cores)},
        "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "index_build_s": t_index,
    }
    print(json.dumps(line), flush=True)


def cxx_extra(args):
    """The drop-in C++ call itself: SeqLib::BWAAligner::alignSequences (UnalignedSequenceVector in, BamRecords out) on a 3 Gb / 24-contig
    reference it builds with BWAIndex::ConstructIndex -- tests/cxx/bench_align_sequences.cpp, run as its own process."""
    exe = os.path.join(ROOT, "tests", "cxx", "bench_align_sequences")
    src = exe + ".cpp"
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "seqlib_b200", "cxx")])
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-o", exe, src, "-L" + os.path.join(ROOT, "seqlib_b200"),
                               "-lSeqLibB200", "-lseqlib_b200", "-pthread", "-Wl,-rpath," + os.path.join(ROOT, "seqlib_b200")])
    ref_mb = max(1, args.ref_len // 1_000_000)
    n = int(os.environ.get("B200_BENCH_CXX_READS", 4_000_000))
    r = subprocess.run([exe, str(ref_mb), str(n)], capture_output=True, text=True, timeout=900)
    if r.returncode != 0:
        raise RuntimeError((r.stdout + r.stderr)[-300:])
    return json.loads(r.stdout.strip().splitlines()[-1])


WAVE_INSTR_PER_CELL = 33.0     # SASS instructions of one wavefront step (66, profiles/r02_sass_extend_wave.txt) / 2 cells per step


def int_peak():
    """Measured integer-ALU issue rate (thread-instructions/s) of the kinds the wavefront is made of:
    profiles/r02_int_peak.json (scripts/microbench/int_peak.cu on this pool's B200s); nominal fallback otherwise."""
    p = os.path.join(ROOT, "profiles", "r02_int_peak.json")
    try:
        rows = [json.loads(l) for l in open(p) if l.strip().startswith("{")]
        alu = [r["thread_instr_per_s"] for r in rows if r["kind"] in ("VIMNMX3.S16x2", "VIADDMNMX.S16x2.RELU", "VIADD.16x2", "PRMT+LOP3")]
        return float(np.median(alu)), "measured: ALU-pipe issue rate (median of VIMNMX3 / VIADDMNMX / VIADD.16x2 / PRMT+LOP3 chains) in profiles/r02_int_peak.json"
    except Exception:
        return 148 * 4 * 16 * 1.965e9, "nominal: 148 SMs x 4 schedulers x 16 ALU lanes x 1.965 GHz"


def ksw_measure(args, n, with_cpu):
    """Config 3: ksw_extend2 n x (150 q, 300 r) microbenchmark through b200_ksw_extend2_batch (the packed 16-bit wavefront kernel,
    then the row-synchronous kernel over whatever it hands back)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from seqlib_b200 import capi
    jobs, qp, tp = cases.c3_tuples_fast(n)
    opt = capi.default_opt()
    mat = np.array(list(opt.mat), dtype=np.int8)
    # the reference's band-trimmed cell count, from the row-synchronous kernel (untimed): the wavefront streams ~2 % more
    os.environ["B200_KSW_WAVE"] = "0"
    out0, cells, _ = capi.ksw_extend2_batch(jobs, qp, tp, mat)
    del os.environ["B200_KSW_WAVE"]
    for _ in range(max(1, args.warmup)):
        capi.ksw_extend2_batch(jobs, qp, tp, mat)
    ms = []
    out = None
    for _ in range(max(1, args.steps)):
        out, _, t = capi.ksw_extend2_batch(jobs, qp, tp, mat)
        ms.append(t)
    t_s = float(np.mean(ms)) * 1e-3
    gcups = cells / t_s / 1e9
    peak, how = int_peak()
    achieved = cells * WAVE_INSTR_PER_CELL / t_s
    res = {"value": gcups, "unit": "GCUPS", "ms_per_step": float(np.mean(ms)), "cells_per_step": int(cells), "pairs": int(n),
           "kernels_agree": bool(out.tobytes() == out0.tobytes()),
           "roofline": {"bound": "int-alu", "kernel": "k_ext_wave (packed s16x2 anti-diagonal wavefront)", "achieved": achieved / 1e9,
                        "peak": peak / 1e9, "unit": "G thread-instr/s", "frac": achieved / peak, "traffic": None,
                        "instr_per_cell": WAVE_INSTR_PER_CELL, "peak_source": how,
                        "note": "cells = the reference's band-trimmed count; instr_per_cell = SASS of the step loop only (pipeline fill/drain, "
                                "row commits and lane idling are the gap between frac and 1)"}}
    if with_cpu:
        from oracle import pyref
        m = min(n, 20000)
        cores = os.cpu_count() or 1
        exp, sec = pyref.ksw_extend2_batch(jobs[:m], qp, tp, mat, n_threads=cores)
        res["parity"] = {"n": int(m), "mismatches": int(np.sum(exp != out[:m])), "against": "oracle/_ref ksw_extend2 (bwa/ksw.c:416-515)"}
        res["cpu_baseline"] = {"value": cells * (m / n) / sec / 1e9, "unit": "GCUPS", "cores": cores, "kind": "reference",
                               "sample": "%d pairs, scalar ksw_extend2 (bwa/ksw.c:416-515) on %d threads" % (m, cores)}
    return res


def run_ksw(args):
    from seqlib_b200 import capi
    capi.set_device(0)
    n = int(os.environ.get("B200_BENCH_KSW_PAIRS", 1_000_000))
    r = ksw_measure(args, n, not args.no_cpu_baseline)
    line = {"metric": "ksw_extend2 GCUPS", "value": r["value"], "unit": "GCUPS", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16x2",
            "data": "synthetic", "config": {"workload": "ksw_extend2 %d x (150q,300r) pairs" % n, "cells_per_step": r["cells_per_step"]},
            "gpu_launches": 2 * args.steps, "roofline": r["roofline"], "kernels_agree": r["kernels_agree"]}
    for k in ("cpu_baseline", "parity"):
        if k in r:
            line[k] = r[k]
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "ksw":
        return run_ksw(args)
    import torch
    import torch.distributed as dist
    from seqlib_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    capi.set_device(local)
    dev = torch.device("cuda", local)
    n_per = args.reads
    L = args.read_len
    opt = capi.default_opt()
    cpu_line = None
    ridx = None
    t_index = t_bcast = 0.0

    # ---- inputs: rank 0 generates reference + all reads, builds the index on its GPU ------------------
    if rank == 0:
        pac, ctg, seqs_all, off_all, t_gen = make_workload(args, n_per * world)
        t0 = time.time()
        want_cpu = (world == 1 and not args.no_cpu_baseline)
        idx = capi.Index.construct_pac(pac, args.ref_len, ctg, keep_host=want_cpu)
        torch.cuda.synchronize()
        t_index = time.time() - t0
        if want_cpu:
            try:
                from oracle import pyref
                cores = os.cpu_count() or 1
                ridx = pyref.RefIndex.from_view(idx.view(), keep=idx)
                v, n_s, t_s = cpu_baseline_sample(ridx, seqs_all, off_all, L, opt, args.cpu_seconds, cores)
                cpu_line = {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference",
                            "sample": "%d of the same reads through mem_process_seqs (bwa/bwamem.c:1235-1264), %d threads, %.1f s" % (n_s, cores, t_s)}
            except Exception as e:  # the baseline is reported, never required for the GPU number
                cpu_line = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    if world > 1:
        # the product's own multi-GPU entry points (include/seqlib_b200.h, csrc/dist.cu): ONE NCCL broadcast of the index image,
        # one scatter of the read batch by contiguous shard.  torch.distributed only carries the 128-byte NCCL id,
        # the barriers and the max-over-ranks reduction of the timings.
        uid = [capi.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = capi.Comm(uid[0], rank, world)
        dist.barrier()
        idx = comm.index_bcast(idx if rank == 0 else None, 0)
        t_bcast = comm.last_bcast_ms() / 1e3
        seqs, sb, se = comm.reads_scatter(seqs_all if rank == 0 else None, n_per * world, L, 0)
        assert se - sb == n_per
        if rank == 0:
            del seqs_all
    else:
        seqs = seqs_all
    off = np.arange(n_per + 1, dtype=np.int64) * L
    from seqlib_b200 import shard
    ids = shard.read_ids(rank * n_per, (rank + 1) * n_per)          # a function of the GLOBAL read index (tests/test_dist_cpu.py)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value leg: reads resident in HBM, all kernels, results left on the device ----------------------
    batch = capi.Batch(idx, (seqs, off), opt, ids)
    for _ in range(args.warmup):
        batch.run()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    launches = 0
    dev_ms = 0.0
    stage = {"ms_seed": 0.0, "ms_chain": 0.0, "ms_extend": 0.0, "ms_finalize": 0.0}
    stats = None
    for _ in range(args.steps):
        launches += batch.run()
        stats = capi.last_stats()
        dev_ms += stats["ms_total"]
        for k in stage:
            stage[k] += stats[k]
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms / 1000.0, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_s, wall_s = float(t[0]), float(t[1])
    n_hits_dev = None
    res = batch.fetch()
    n_hits_dev = len(res.hits)
    mapped = float((np.diff(res.hit_off) > 0).mean())
    par = None
    if rank == 0 and world == 1:
        par = parity_block(res, seqs, off, ids, ctg, ridx, opt, args.parity_reads, os.cpu_count() or 1)
        ridx = None
    del res
    batch.close()

    # ---- e2e leg: the C ABI call on pinned host buffers, H2D of reads and D2H of results inside the timed region
    pin_seqs = torch.from_numpy(seqs).pin_memory()
    pin_off = torch.from_numpy(off).pin_memory()
    pin_ids = torch.from_numpy(ids).pin_memory()
    seqs_p, off_p, ids_p = pin_seqs.numpy(), pin_off.numpy(), pin_ids.numpy()
    h = capi.align_raw(idx, seqs_p, off_p, opt, ids_p)
    summ = capi.results_summary(h)
    capi.results_free(h)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps)            # the same K steps as the device-resident leg
    for _ in range(e2e_steps):
        h = capi.align_raw(idx, seqs_p, off_p, opt, ids_p)
        capi.results_free(h)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te[0])
    h2d = int(seqs.nbytes + off.nbytes + ids.nbytes)
    d2h = int(summ["n_hits"] * 144 + summ["n_cigar"] * 4 + summ["n_md"] + (n_per + 1) * 8)

    if rank == 0:
        peaks, how = measured_peaks()
        total_reads = n_per * world * args.steps
        value = total_reads / dev_s
        # Rooflines.  The stage that takes longest is the banded extension (k_extend_wave, packed 16-bit anti-diagonal wavefront): an
        # integer recurrence bound by the ALU pipe, reported as `roofline`.  The seeding stage (FM-index gathers) is the HBM-side
        # kernel the metric names, reported as `roofline_hbm`: algorithmic bytes = 32 B per Occ block fetched + 32 B per prefix-chain
        # entry + 48 B per text-path request (one suffix-array sector or two text sectors) + the read bases.
        seed_s = stage["ms_seed"] / 1000.0 / args.steps
        ext_s = stage["ms_extend"] / 1000.0 / args.steps
        occ_per_step = stats["occ_blocks"]
        chain_ent = stats.get("tab_lookups_hi", 0)
        text_req = stats.get("tab_lookups_lo", 0)
        chunk = min(int(os.environ.get("B200_CHUNK", 1 << 21)), n_per)
        n_seed_launches = max(1, len(chunk_plan(n_per, chunk)))
        alg_bytes = occ_per_step * 32 + chain_ent * 32 + text_req * 48 + n_per * L
        achieved = alg_bytes / seed_s / 1e9 if seed_s > 0 else 0.0
        tpr, tsrc = seed_traffic_per_read()
        int_pk, int_src = int_peak()
        int_pk /= 1e9
        ext_ach = stats["sw_cells"] * WAVE_INSTR_PER_CELL / ext_s / 1e9 if ext_s > 0 else 0.0
        line = {
            "metric": "150bp reads/sec (seed+chain+SW end-to-end)", "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dev_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16x2 (extension) / int32, u64 (seeding, chaining, CIGAR)", "data": "synthetic",
            "config": {"workload": "%d x %dbp synthetic reads per GPU vs %d bp random reference (24 contigs), 1%% substitutions" % (n_per, L, args.ref_len),
                       "reads_per_gpu_per_step": n_per, "read_len": L, "ref_len": args.ref_len,
                       "parallelism": "reads sharded x%d, index replicated (one NCCL broadcast)" % world,
                       "l2": "inputs larger than L2 (index %.1f GB, reads %.2f GB per GPU)" % (idx.blob_bytes() / 1e9, n_per * L / 1e9)},
            "e2e": {"value": n_per * world * e2e_steps / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "int-alu", "kernel": "k_extend_wave<4> (ksw_extend2 as a packed s16x2 anti-diagonal wavefront), one launch per chunk; the longest stage of the step",
                         "achieved": ext_ach, "peak": int_pk, "unit": "G thread-instr/s", "frac": ext_ach / int_pk,
                         "traffic": 1.591e9 * (n_per / n_seed_launches) / 1e6, "traffic_source": "profiles/r02_ncu_full_extend_wave_v2.txt: dram read 1.185 GB + write 0.406 GB per 10^6 reads (ncu --set full)",
                         "peak_source": int_src,
                         "launches_per_step": n_seed_launches, "kernel_ms_per_launch": 1000.0 * ext_s / n_seed_launches, "kernel_ms": 1000.0 * ext_s,
                         "cells_per_launch": stats["sw_cells"] / n_seed_launches, "gcups": stats["sw_cells"] / ext_s / 1e9 if ext_s > 0 else 0.0,
                         "instr_per_cell": WAVE_INSTR_PER_CELL,
                         "note": "algorithmic work = band cells of the reference's ksw_extend2 (bwa/ksw.c:416-515) x the SASS instructions of one wavefront step per cell; "
                                 "the gap to 1 is pipeline fill / drain of the 8-row blocks, row commits and lanes whose group finished its block early (19 of 32 lanes per instruction)"},
            "roofline_hbm": {"bound": "hbm", "kernel": "seeding stage: k_seed2 (SMEM passes: prefix-chain table + Occ blocks + text path) + k_seed3 (bwt_seed_strategy1), one pair of launches per chunk",
                         "achieved": achieved, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                         "frac": achieved / peaks.get("hbm_gbs", 6650.0), "traffic": (tpr * n_per / n_seed_launches if tpr else None), "peak_source": how,
                         "traffic_source": tsrc, "algorithmic_bytes_per_launch": alg_bytes / n_seed_launches,
                         "launches_per_step": n_seed_launches, "kernel_ms_per_launch": 1000.0 * seed_s / n_seed_launches,
                         "algorithmic_bytes_per_read": alg_bytes / n_per, "kernel_ms": 1000.0 * seed_s,
                         "gathers_per_read": {"occ_blocks_32B": occ_per_step / n_per, "chain_entries_32B": chain_ent / n_per, "text_path_requests_48B": text_req / n_per},
                         "note": "dependent random gathers.  Round 1 fetched 1391 Occ blocks per read (44.7 KB); the prefix-chain table (one 32-byte entry answers a forward prefix or a whole backward row) "
                                 "and the text path (size-one intervals are followed through the text) leave ~240 gathers / ~8 KB per read, so the ALGORITHMIC GB/s falls while the stage got 3.5x faster: "
                                 "the kernel is no longer bound by HBM but by dependent-gather latency and instruction issue (ncu: issue slots 55 % busy, 9 of 32 lanes per instruction, DRAM 25 % of peak, "
                                 "every L2 miss fills a 128-byte line: profiles/r02_ncu_full_seed2_mask_v0.txt)"},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage.items()},
            "wall_s_timed_region": wall_s, "mapped_fraction": mapped, "hits_per_step": n_hits_dev,
            "index_build_s": t_index, "index_bcast_s": t_bcast, "spill_reads_per_step": stats["n_overflow"],
            "sw_cells_per_step": stats["sw_cells"], "occ_blocks_per_read": occ_per_step / n_per,
            "seed_gathers_per_read": {"occ_blocks": occ_per_step / n_per, "chain_entries": chain_ent / n_per, "text_path_requests": text_req / n_per},
        }
        if cpu_line is not None:
            line["cpu_baseline"] = cpu_line
        if par is not None:
            line["parity"] = par
        if world == 1 and not args.no_extra:
            try:
                idx.close()
                line["extra"] = {"config4_fermi_assemble": fermi_extra(args, not args.no_cpu_baseline)}
            except Exception as e:   # secondary numbers never fail the headline line
                line["extra"] = {"config4_fermi_assemble": {"error": str(e)}}
            try:
                line["extra"]["cxx_alignSequences"] = cxx_extra(args)
            except Exception as e:
                line["extra"]["cxx_alignSequences"] = {"error": str(e)}
            try:
                line["extra"]["config3_ksw"] = ksw_measure(args, int(os.environ.get("B200_BENCH_KSW_PAIRS", 1_000_000)), not args.no_cpu_baseline)
            except Exception as e:
                line["extra"]["config3_ksw"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
