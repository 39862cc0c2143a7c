"""DRAM traffic per launch from `ncu --set full` reports -> profiles/traffic.json (read by bench.py's roofline).
Usage: python tools/ncu_traffic.py KEY::report.ncu-rep[:kernel-substring[+kernel-substring...]] ...
  e.g. "bgnn_gatv2_bwd_f32[c=64]::gpurun_out/r01m_gat.ncu-rep:gatv2_bwd_dst+gatv2_bwd_src"
Every kernel substring must match exactly one profiled launch (the first match is taken); bytes are summed."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    u = unit.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}[u]


def main():
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        table = json.load(open(path))
    except Exception:
        table = {}
    for arg in sys.argv[1:]:
        key, spec = arg.split("::", 1)
        rep, _, names = spec.partition(":")
        hdr, units, rows = rows_of(rep)
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        ti = hdr.index("gpu__time_duration.sum")
        total, used, ms = 0.0, [], 0.0
        for sub in (names.split("+") if names else [""]):
            m = [r for r in rows if sub in r[ki]]
            if not m:
                raise SystemExit("no launch matching %r in %s" % (sub, rep))
            r = m[0]
            total += to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
            ms += float(r[ti].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[ti].lower()]
            used.append(r[ki].split("(")[0])
        table[key] = {"bytes": total, "kernels": used, "ncu_ms": ms,
                      "source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum (%s)" % os.path.basename(rep)}
        print(key, "%.3f GB over %s" % (total / 1e9, used))
    json.dump(table, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
