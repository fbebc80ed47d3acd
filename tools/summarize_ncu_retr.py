"""gpurun_out/prof_retr.ncu-rep (+ retr_launches.csv) -> profiles/<tag>_ncu_retrieval_tc.json / _launch_list.md."""
import csv, json, subprocess, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = REPO / "gpurun_out" / "prof_retr.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__sass_inst_executed_op_utcmma.sum", "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")].replace("spn::<unnamed>::", "")[:80]}
    for k in KEYS:
        if k in hdr:
            d[k] = r[hdr.index(k)] + " " + units[hdr.index(k)]
    st = sorted(((float(r[hdr.index(h)].replace(",", "")), h) for h in stall), reverse=True)[:4]
    d["top_stalls_per_issue"] = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): v for v, h in st}
    out.append(d)
summary = {"source": "ncu --set full --clock-control none --import-source on -k regex:'row_|sim_tc_gemm|target_score_warp|rank_prep|rank_resolve|split_rows' "
                     "(python tools/prof_retr.py), C2: 4181 queries x 2297 gallery x 640-d; launches in order: "
                     "top-50 exact values, top-50 index-exact, rank-of-target (+ their sweeps)",
           "kernels": out}
(REPO / "profiles").mkdir(exist_ok=True)
(REPO / "profiles" / f"{tag}_ncu_retrieval_tc.json").write_text(json.dumps(summary, indent=1))
ll = REPO / "gpurun_out" / "retr_launches.csv"
if ll.exists():
    lines = [l for l in ll.read_text().splitlines() if not l.startswith("==")]
    rr = list(csv.reader(lines))
    h = rr[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    seq = []
    for r in rr[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("spn::<unnamed>::", "")
        seq.append((name, float(r[vi].replace(",", ""))))
    # the last iteration of tools/prof_retr.py: top-50 (exact values), top-50 (index-exact), rank
    md = ["# ncu launch list, tensor-core retrieval at C2 (gpu__time_duration.sum, --clock-control none --cache-control none)",
          "", "`python tools/prof_retr.py`: per iteration sim_topk(K=50, exact values), sim_topk(K=50, index-exact), sim_rank(+group).",
          "Serialised launches: compare shares, not absolutes.", "", "## last iteration, in launch order", ""]
    tail = seq[-16:]
    md += [f"- `{n[:80]}` {v / 1e3:.2f} us" for n, v in tail]
    (REPO / "profiles" / f"{tag}_ncu_retrieval_launch_list.md").write_text("\n".join(md) + "\n")
print(json.dumps(summary)[:300])
