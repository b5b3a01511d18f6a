import sys
sys.path.insert(0, '/root/repo')
from cornelis_b200 import binding, scenes
sc = binding.Scene(scenes.cornell_box())
for (w,h,spp) in [(512,512,64),(512,512,16),(128,128,64),(1920,1080,16),(1920,1080,256)]:
    sc.render_accumulate(w,h,spp)
    best = min(sc.render_accumulate(w,h,spp)["gpu_ms"] for _ in range(5))
    print(w,h,spp, f"{best:.3f} ms", f"{w*h*spp/best/1e3:.0f} Msamples/s")
