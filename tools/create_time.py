"""cost of Mapper.create() / free of managed maps at nside 4096 (measured: 9 + 17 ms to create a POS + SHE pair, 86 ms to free it -- cudaFree waits for the device)"""
import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import heracles_b200 as hb
m = hb.CudaHealpixMapper(4096, 8192, deconvolve=False, sync=False, pixel_weights=None)
m.create()  # context warm-up
for rep in range(3):
    t0 = time.perf_counter(); a = m.create(spin=0); t1 = time.perf_counter(); b = m.create(2, spin=2); t2 = time.perf_counter()
    a /= 2.0; m.context.synchronize(); t3 = time.perf_counter()
    del a, b; t4 = time.perf_counter()
    print(f"create pos {1e3*(t1-t0):.1f} ms, she {1e3*(t2-t1):.1f} ms, divide {1e3*(t3-t2):.1f} ms, free both {1e3*(t4-t3):.1f} ms")
