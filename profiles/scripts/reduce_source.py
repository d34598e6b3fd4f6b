"""Reduce an `ncu --page source --csv` dump to the table committed under profiles/: per-reason totals + the top-N SASS
instructions by warp-stall samples.   python profiles/scripts/reduce_source.py in_source.csv out_stalls.csv [N]"""
import csv
import sys

src, dst = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(src)))
kernel = rows[0][1]
hdr = rows[1]
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ix = {h: hdr.index(h) for h in hdr}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = {k: sum(int(r[ix[k]] or 0) for r in body) for k in reasons}
total = sum(tot.values())
with open(dst, "w") as f:
    f.write(f"# {kernel}\n")
    f.write(f"# total warp-stall samples {total}; by reason: " + ", ".join(
        f"{k[6:]}={v} ({100.0 * v / total:.1f}%)" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 50 > total) + "\n")
    f.write("sass_index,instruction,samples,share,instructions_executed,top_reasons\n")
    ranked = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top_n]
    for i in ranked:
        r = body[i]
        n = int(r[ix["# Samples"]] or 0)
        top = sorted(((int(r[ix[k]] or 0), k[6:]) for k in reasons), reverse=True)[:2]
        f.write(f'{i},"{r[ix["Source"]].strip()}",{n},{n / total:.4f},{r[ix["Instructions Executed"]]},'
                + ";".join(f"{k}={v}" for v, k in top if v) + "\n")
print("wrote", dst)
