#!/usr/bin/env python
"""The pipes that bound the fused solve kernel, from an ncu report: FP64 (DFMA + DMMA share one pipe),
the L1TEX LSU data pipe (shared-memory loads / stores, shuffles, global accesses: one wavefront per cycle per SM),
issue slots; per-instance instruction and wavefront counts.

    python tools/ncu_lsu.py report.ncu-rep [instances]"""
import csv
import io
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "kernel time [us]"),
        ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe busy (DFMA+DMMA) [%]"),
        ("smsp__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "  of which DMMA [%]"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data pipe busy [%]"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", "  shared loads [%]"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum.pct_of_peak_sustained_elapsed", "  shared stores [%]"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "  shared total incl. shuffles [%]"),
        ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum.pct_of_peak_sustained_elapsed", "  global loads [%]"),
        ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum.pct_of_peak_sustained_elapsed", "  global stores [%]"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy [%]"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active [% of 64]"),
        ("launch__registers_per_thread", "registers / thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank-conflict wavefronts"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written")]
PER = [("smsp__inst_executed.sum", "warp instructions"), ("smsp__inst_executed_pipe_fp64.sum", "FP64-pipe instructions (DFMA/DADD/DMUL/DSETP)"),
       ("sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "DMMA"), ("smsp__sass_inst_executed_op_shared_ld.sum", "LDS"),
       ("smsp__sass_inst_executed_op_shared_st.sum", "STS"), ("smsp__sass_inst_executed_op_global_ld.sum", "LDG"),
       ("smsp__sass_inst_executed_op_global_st.sum", "STG"), ("smsp__sass_inst_executed_op_local_ld.sum", "LDL (spill)"),
       ("smsp__sass_inst_executed_op_local_st.sum", "STL (spill)"),
       ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "shared-load wavefronts"),
       ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "shared-store wavefronts"),
       ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts incl. shuffles"),
       ("l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum", "global-load wavefronts")]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
    inst = float(sys.argv[2]) if len(sys.argv) > 2 else None
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("==", d.get("Kernel Name", "?")[:100])
        for k, name in KEYS:
            if k in d:
                print(f"  {name:42s} {float(d[k].replace(',', '')):14.3f} {units[hdr.index(k)]}")
        if inst:
            print(f"  per instance (= per warp; {int(inst)} instances):")
            for k, name in PER:
                if k in d:
                    print(f"    {name:52s} {float(d[k].replace(',', '')) / inst:10.1f}")
        st = sorted(((float(v), k[len(STALL):-len('_per_issue_active.ratio')]) for k, v in d.items()
                     if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v), reverse=True)
        print("  stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:8]))


if __name__ == "__main__":
    main()
