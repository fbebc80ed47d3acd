"""profiles/<tag>_sass_evidence.md: per-kernel counts of the Blackwell-specific SASS mnemonics in libspn_b200.so."""
import re, subprocess, sys
from collections import OrderedDict
from pathlib import Path
REPO = Path(__file__).resolve().parent.parent
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
sass = subprocess.run(["cuobjdump", "-sass", str(REPO / "spn4cir_b200" / "libspn_b200.so")], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
COLS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "SHFL", "ATOMS", "RED", "ATOMG", "ST.E", "LD.E", "MEMBAR"]
kern = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        kern[cur] = {c: 0 for c in COLS}
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for c in COLS:
        if op == c or op.startswith(c + ".") or (c in ("ST.E", "LD.E") and op.startswith(c)):
            kern[cur][c] += 1
rows = sorted(kern.items(), key=lambda kv: -(kv[1]["UTCHMMA"] * 1000 + kv[1]["UTMALDG"] + kv[1]["UTMAREDG"]))
out = [f"# SASS evidence (cuobjdump -sass spn4cir_b200/libspn_b200.so, sm_100a): tcgen05 / TMA / TMEM instruction counts per kernel", "",
       "UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce (UTMAREDG in bank_ce_kernel<2, *, *, 2> is the "
       "reduce-add epilogue that targets the home rank's peer-mapped window over NVLink when the bank is row-sharded), LDTM / STTM = "
       "tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; ST.E / LD.E in shard_home_prologue_kernel and RED / ATOMG in "
       "bank_ce_kernel are the generic stores / loads / reductions that go to peer addresses in the NCCL-free exchange; MEMBAR.SC.SYS in "
       "peer_barrier_kernel is the system-scope fence of the device-side barrier.", "",
       "| kernel | " + " | ".join(COLS) + " |", "|---|" + "---|" * len(COLS)]
for name, c in rows:
    short = name.replace("spn::(anonymous namespace)::", "").replace("void ", "")
    short = re.sub(r"\(.*", "", short)[:70]
    out.append(f"| `{short}` | " + " | ".join(str(c[k]) for k in COLS) + " |")
tot = {k: sum(c[k] for c in kern.values()) for k in COLS}
out.append("| **total** | " + " | ".join(str(tot[k]) for k in COLS) + " |")
(REPO / "profiles" / f"{tag}_sass_evidence.md").write_text("\n".join(out) + "\n")
print("\n".join(out[-12:]))
