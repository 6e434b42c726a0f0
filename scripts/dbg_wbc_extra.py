"""Development aid: the config-4 extra of bench.py on its own (device rate, end to end, end to end from robot states)."""
import argparse, json, sys
sys.path.insert(0, ".")
import bench

args = argparse.Namespace(no_cpu_baseline=True, cpu_budget=0.0)
out = bench.extra_config4(args, 0, 0, 1, steps=int(sys.argv[1]) if len(sys.argv) > 1 else 10)
print(json.dumps({k: out[k] for k in ("value", "e2e", "converged_fraction", "mean_iterations", "e2e_scene_on_device")}))
