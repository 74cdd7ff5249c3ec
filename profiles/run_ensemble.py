"""One launch of the engine-2 x-line kernel for ncu captures:
    ncu --set full --clock-control none --import-source on -k regex:k_problem_xline -s 1 -c 1 -o gpurun_out/xline_r02 \
        python profiles/run_ensemble.py 592 2
members (default 592 = 4 per SM) x steps (default 2) of the BASELINE config-4 ensemble; prints cell-steps and cell-iterations
of the profiled call so that per-cell figures can be formed from the counters."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

members = int(sys.argv[1]) if len(sys.argv) > 1 else 592
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
bench.NB_TOTAL = 4096
eng, _ = bench.make_ensemble(0, members, 0)
eng.step(steps)            # launch 0 (skipped by -s 1): first-step transient, issue order
st = eng.step(steps)       # launch 1: the profiled one
print({"members": members, "steps": steps, "cell_steps": members * 6400 * steps, "cell_iterations": st["iters_sum_all"] * 6400,
       "ms": 1e3 * st["seconds_device"]})
eng.close()
