import sys, json
for l in sys.stdin:
    r=json.loads(l)
    print(r["op"].ljust(32), f'{r["ms"]:9.3f} ms  {r["achieved_gbs"]:8.1f} GB/s  frac {r["frac_of_hbm"]:.3f}  rate {r["rate"]:.3e} {r["unit"]}  cpu {r["cpu_rate"]}')
