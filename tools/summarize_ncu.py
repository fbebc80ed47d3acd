"""Turn gpurun_out/prof_bank_ce.ncu-rep (+ launches.csv) into the tracked summaries under profiles/."""
import csv, json, subprocess, sys
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
# usage: summarize_ncu.py <tag> [report.ncu-rep] [out name] [launches.csv] [description]
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = Path(sys.argv[2]) if len(sys.argv) > 2 else REPO / "gpurun_out" / "prof_bank_ce.ncu-rep"
out_name = sys.argv[3] if len(sys.argv) > 3 else "ncu_bank_ce_fused"
ll_path = Path(sys.argv[4]) if len(sys.argv) > 4 else REPO / "gpurun_out" / "launches.csv"
desc = sys.argv[5] if len(sys.argv) > 5 else "C3: B=512, N=100000, D=640"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__sass_inst_executed_op_utcmma.sum", "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "smsp__sass_inst_executed_op_tmem_stt.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum", "gpc__cycles_elapsed.max.per_second"]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            d[k] = {"value": r[i], "unit": units[i]}
    out.append(d)
def num(d, k):
    try: return float(d[k]["value"].replace(",", ""))
    except Exception: return None
def to_bytes(d, k):
    v = num(d, k)
    if v is None: return None
    u = d[k]["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
main = out[-1]
summary = {"source": "ncu --set full --clock-control none --import-source on -k regex:bank_ce_kernel (python bench.py --profile --steps 1), " + desc,
           "kernels": out,
           "dram_bytes_per_launch": (to_bytes(main, "dram__bytes_read.sum") or 0) + (to_bytes(main, "dram__bytes_write.sum") or 0),
           "l2_to_sm_bytes_per_launch": to_bytes(main, "l1tex__m_xbar2l1tex_read_bytes.sum")}
(REPO / "profiles").mkdir(exist_ok=True)
(REPO / "profiles" / f"{tag}_{out_name}.json").write_text(json.dumps(summary, indent=1))
# launch list
ll = ll_path
if ll.exists():
    lines = [l for l in ll.read_text().splitlines() if not l.startswith("==")]
    rr = list(csv.reader(lines))
    h = rr[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    tot = {}
    seq = []
    for r in rr[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("spn::<unnamed>::", "")
        ns = float(r[vi].replace(",", ""))
        seq.append((name, ns)); tot[name] = tot.get(name, 0) + ns
    total = sum(tot.values())
    md = ["# ncu launch list (gpu__time_duration.sum, --clock-control none) -- python bench.py --profile --steps 2", "",
          "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "", "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for n, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        md.append(f"| `{n[:90]}` | {sum(1 for s in seq if s[0] == n)} | {v / 1e3:.1f} | {100 * v / total:.1f} % |")
    md += ["", "## in launch order (last step)", ""] + [f"- `{n[:90]}` {v / 1e3:.2f} us" for n, v in seq[-12:]]
    (REPO / "profiles" / f"{tag}_{out_name}_launch_list.md").write_text("\n".join(md) + "\n")
print(json.dumps({k: v for k, v in summary.items() if k != "kernels"}))
for d in out:
    print(d["kernel"][:80], {k.split(".")[0]: v["value"] for k, v in d.items() if k != "kernel"})
