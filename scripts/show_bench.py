"""Dev tool: one-line digest of a bench.py JSON line (file may contain other lines)."""
import json, sys
for ln in open(sys.argv[1]):
    ln = ln.strip()
    if not ln.startswith("{"):
        continue
    d = json.loads(ln); r = d.get("roofline", {}); inf = r.get("inference_forward", {})
    print(f"n_gpus {d['n_gpus']}  value {d['value']/1e6:.1f} M/s  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']/1e6:.1f} M/s  "
          f"kernel {r.get('kernel_ms_avg', 0)*1e3:.1f} us frac {r.get('frac', 0):.3f} [{r.get('kernel','')[:24]}]  "
          f"bwd {(r.get('backward_kernel_ms_avg') or 0)*1e3:.1f} us  inf {inf.get('ms_median', 0)*1e3:.1f} us frac {inf.get('frac', 0):.3f}  "
          f"launches {d.get('gpu_launches')}  clocks {d.get('clocks')}  loss {d.get('final_loss')}")
    if "cpu_baseline" in d:
        c = d["cpu_baseline"]; print(f"   cpu_baseline {c['value']:.0f} samples/s on {c['cores']} cores ({c['kind']})")
