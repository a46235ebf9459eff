#!/usr/bin/env python
"""Summarise an ncu report of colour_sweep_fast into markdown (profiles/*.md).

    python tools/ncu_summary.py gpurun_out/prof_final_r1.ncu-rep "<command that produced it>" rows sweeps > profiles/...

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (no GPU needed).
rows/sweeps: replica rows and sweeps of the captured launch (to normalise per word / per attempt).
"""
import collections
import csv
import subprocess
import sys

NSPINS, LANES = 65536, 64

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
       "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main():
    rep, cmd, rows, sweeps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    words = float(NSPINS) * rows * sweeps
    raw = page(rep, "raw")
    hdr, units, val = raw[0], raw[1], raw[2]
    get = {h: (v, u) for h, u, v in zip(hdr, units, val)}
    kname = get.get("Kernel Name", ("?", ""))[0]
    print("# ncu summary: `%s`\n" % kname)
    print("Command (one B200, under gpurun): `%s`\n" % cmd)
    print("The captured launch: %d natural-order sweeps of the 256x256, P=64, R=%d workload "
          "(%.3g words = %.3g attempts) in ONE dataflow launch.  Times under ncu are serialised/cold; use shares "
          "and per-word counts.\n" % (sweeps, rows, words, words * LANES))
    print("| metric | value | unit |\n|---|---|---|")
    for m in RAW:
        if m in get:
            print("| %s | %s | %s |" % (m, get[m][0], get[m][1]))
    rd = to_bytes(*get["dram__bytes_read.sum"])
    wr = to_bytes(*get["dram__bytes_write.sum"])
    inst = float(get["smsp__inst_executed.sum"][0])
    alg = 2 * 8 * words
    print("\nDerived:\n")
    print("* DRAM traffic per launch: %.3f GB read + %.3f GB written = %.3f GB; algorithmic bytes (each packed "
          "word read once, written once) = %.3f GB -> traffic / algorithmic = %.3f (neighbour words are L2 hits)."
          % (rd / 1e9, wr / 1e9, (rd + wr) / 1e9, alg / 1e9, (rd + wr) / alg))
    print("* warp-instructions per 32 words (2048 attempts): %.0f; per attempt: %.4f."
          % (inst / (words / 32), inst / (words * LANES)))
    src = page(rep, "source")
    h = src[1]
    ie, si, ss = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    stalls = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
    tot, ops = collections.Counter(), collections.Counter()
    for r in src[2:]:
        if len(r) <= ie:
            continue
        for i, n in stalls:
            tot[n] += int(r[i] or 0)
        op = r[si].strip().split()
        if op and op[0].startswith("@"):
            op = op[1:]
        if op:
            ops[op[0].split(".")[0]] += int(r[ie])
    T = float(sum(tot.values()))
    print("\nWarp stall samples (all instructions): " +
          ", ".join("%s %.1f%%" % (n.replace("stall_", ""), 100 * v / T) for n, v in tot.most_common(9)))
    print("\nInstruction mix (executed warp-instructions per 32 words):\n\n| opcode | per 32 words |\n|---|---|")
    for o, v in ops.most_common(16):
        print("| %s | %.1f |" % (o, v / (words / 32)))


if __name__ == "__main__":
    main()
