#!/usr/bin/env python
"""TraceBench — benchmark and memory profiler for the `tracer` program.

Same command line, same matrix and the same benchmark.txt format as the reference's test_automation.py
(its :18-20 matrix: depth 10 x samples {40,80,200,400,1000,2000,5000,10000} x {spheres,maze,mirrors};
`-o/--one` runs the single case given by -s/-d/-p; a timed-out case writes `<id>;DNF;DNF;DNF` and skips the larger
sample counts of that scene, :32-36,116-125).  `tracer` itself appends `<id>;HH:MM:SS.ms;`; this driver appends
`<cpuMiB>;<gpuMiB>\\n` (:103-112).  Differences: GPU memory is polled through NVML (pynvml) every 100 ms instead of
parsing `nvidia-smi -lms=500` text, and the tracer binary / working directory can be chosen.
"""
import argparse
import os
import resource
import subprocess
import threading
import time

BENCHMARK_FILE = "benchmark.txt"
TIMEOUT = 86400
DEPTHS = [10]
SAMPLES = [40, 80, 200, 400, 1000, 2000, 5000, 10000]
SCENES = ["scenes/spheres.json", "scenes/maze.json", "scenes/mirrors.json"]


class GpuMemoryPoller(threading.Thread):
    """Peak device memory used by the child process (MiB), via NVML's per-process accounting."""

    def __init__(self, pid_getter):
        super().__init__(daemon=True)
        self.pid_getter, self.peak, self.stop_flag = pid_getter, 0.0, False

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handles = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in range(pynvml.nvmlDeviceGetCount())]
        except Exception:
            return
        while not self.stop_flag:
            pid = self.pid_getter()
            for h in handles:
                try:
                    for p in pynvml.nvmlDeviceGetComputeRunningProcesses(h):
                        if pid is None or p.pid == pid:
                            self.peak = max(self.peak, (p.usedGpuMemory or 0) / (1024 * 1024))
                except Exception:
                    pass
            time.sleep(0.1)


def write_timeout(scene, sample, depth):
    ident = os.path.splitext(os.path.basename(scene))[0] + "D" + str(depth) + "S" + str(sample)
    with open(BENCHMARK_FILE, "a") as f:
        f.write(ident + ";DNF;DNF;DNF\n")


def run_test(exe, scene, depth, sample):
    """Returns True if the case timed out (the caller then skips the larger sample counts)."""
    print(f"Starting: {scene} Depth={depth} Samples={sample}")
    proc = subprocess.Popen([exe, f"-d={depth}", f"-s={sample}", scene])
    poller = GpuMemoryPoller(lambda: proc.pid)
    poller.start()
    timed_out = False
    try:
        proc.wait(timeout=TIMEOUT)
    except subprocess.TimeoutExpired:
        print("\nTimeout! Skipping further execution for scene/depth combination.\n")
        proc.kill()
        proc.wait()
        timed_out = True
    poller.stop_flag = True
    poller.join(timeout=1)
    if timed_out:
        write_timeout(scene, sample, depth)
        return True
    cpu_mib = str(round(resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss / 1024, 2))
    print("CPU Memory used: " + cpu_mib + " MiB")
    print("GPU Memory used: " + str(round(poller.peak, 1)) + " MiB\n")
    with open(BENCHMARK_FILE, "a") as f:
        f.write(cpu_mib + ";" + str(round(poller.peak, 1)) + "\n")
    return False


def main():
    ap = argparse.ArgumentParser(prog="TraceBench", description="Benchamrk tool and memory profiler for tracer program.")
    ap.add_argument("-o", "--one", action="store_false", help="Enable execution of single test case.")
    ap.add_argument("-s", "--samples", default="40", help="Specifies number of samples per pixel.")
    ap.add_argument("-d", "--depth", default="10", help="Specifies max number of reflections per ray.")
    ap.add_argument("-p", "--path", default="scenes/spheres.json", help="Specifies path to json file with scene data.")
    ap.add_argument("--tracer", default="./tracer", help="tracer executable (default ./tracer, as the reference expects)")
    args = ap.parse_args()
    if not os.path.exists(args.tracer):
        print("Executable not found")
        return 1
    if os.path.exists(BENCHMARK_FILE):
        os.remove(BENCHMARK_FILE)
    if not args.one:
        run_test(args.tracer, args.path, args.depth, args.samples)
    else:
        for scene in SCENES:
            for depth in DEPTHS:
                too_long = False
                for sample in SAMPLES:
                    if not too_long:
                        too_long = run_test(args.tracer, scene, depth, sample)
                    else:
                        write_timeout(scene, sample, depth)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
