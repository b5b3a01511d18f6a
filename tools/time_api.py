"""Where does the end-to-end overhead go?  Times the C-ABI calls of one e2e step."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
from cornelis_b200 import binding, scenes
flat = scenes.cornell_box(aspect=1080 / 1920)
binding.Scene(flat).close()
for it in range(3):
    t0 = time.perf_counter(); sc = binding.Scene(flat); t1 = time.perf_counter()
    st = sc.render_accumulate(1920, 1080, 64); t2 = time.perf_counter()
    img = sc.resolve(64); t3 = time.perf_counter()
    sc.close(); t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  render(64spp) {1e3*(t2-t1):.2f} ms (gpu {st['gpu_ms']:.2f})  resolve+d2h {1e3*(t3-t2):.2f} ms  destroy {1e3*(t4-t3):.2f} ms")
