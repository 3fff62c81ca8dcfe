"""CPU reference timings for the region-based configs on this box (oracle/_ref = the unmodified reference, 1 thread: the greedy
generator is serial upstream and par_unseq is serial without TBB).  Prints wall times."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import pyoracle
R = pyoracle.load("reference" if pyoracle.available("reference") else "port")
print("checker:", R.kind)
t = time.perf_counter(); b, _ = R.adaptive_iterations("smooth_edge2", "boole_simpson", "size_relative", 1000000, [512, 512], [0, 0], [1, 1]); dt = time.perf_counter() - t
print(f"C3 reference: 1e6 iterations + region->bin 512x512: {dt:.2f} s, mean {b.mean():.6f}")
w = int(sys.argv[1]) if len(sys.argv) > 1 else 128
t = time.perf_counter(); b, _ = R.crespo2021("shade5_64", 65536, 64, 0, [w, w], [0] * 5, [1] * 5); dt = time.perf_counter() - t
print(f"C4 reference: crespo2021(65536, 64) at {w}x{w} bins: {dt:.1f} s = {w*w*64/dt/1e6:.3f} M evals/s, mean {b.mean():.5f}; full 1024x1024 extrapolates to ~{dt*(1024/w)**2:.0f} s")
if len(sys.argv) > 2:      # new rows of round 1b
    tol = float(sys.argv[2])
    t = time.perf_counter(); b, n, _ = R.adaptive_tolerance("smooth_edge2", "boole_simpson", "default_absolute", tol, [512, 512], [0, 0], [1, 1]); dt = time.perf_counter() - t
    print(f"TOL reference: tolerance {tol:g}: {n} leaves + region->bin 512x512: {dt:.2f} s ({n/dt/1e6:.3f} M leaves/s), mean {b.mean():.6f}")
    wf = 64
    t = time.perf_counter(); b = R.crespo2021_infinite("walk", 2, 16384, 16, 64, 0, [wf, wf]); dt = time.perf_counter() - t
    print(f"FUB reference: crespo2021_infinite<2>(16384,16,64) at {wf}x{wf} bins: {dt:.1f} s = {wf*wf*64/dt/1e6:.3f} M residual paths/s, mean {b.mean():.5f}")
