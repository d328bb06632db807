"""prints the few numbers of a bench.py JSON line that tuning runs compare"""
import json
import sys
for path in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(path) if l.startswith("{")][-1])
        r = d.get("roofline") or {}
        e = d.get("e2e") or {}
        print(f"{path}: value={d['value']:.1f} {d['unit']} ms/step={d['ms_per_step']:.2f} e2e={e.get('value')} "
              f"frac={r.get('frac')} busy={r.get('fp64_pipe_busy_model')} kernel={r.get('kernel')} clocks={d.get('clocks')}")
    except Exception as ex:  # noqa: BLE001
        print(path, "unreadable:", ex)
