import json, sys
rows = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/configs_r01.json"))
for d in rows:
    if "solve_ms" in d:
        print(f"{d['config']:48s} k1 {d['k1_GBps']:7.0f} GB/s  solve {d['solve_ms']:8.4f} ms  {d['solves_per_s']:12.0f} /s  it {d['mean_iterations']:.2f}")
    else:
        print(f"{d['config']:48s} {d['ms']:.4f} ms  {d['GBps']:.0f} GB/s")
