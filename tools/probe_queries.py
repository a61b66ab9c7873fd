import importlib, sys, time, numpy as np
sys.path.insert(0,'.')
cq = importlib.import_module('swift-game-engine_b200')
sc = cq.scenes
parts = sc.mirror_scene(use_hulls=False)
g = cq.CollisionQuery(parts)
lo, hi = sc.scene_aabb(parts[1:])
n = 1<<20
rays = sc.gen_rays(n, lo, hi, seed=1, expand=2.0)
caps = sc.gen_capsules(n, lo, hi, seed=2)
casts = sc.gen_casts(n, lo, hi, seed=3)
for name, fn in (('raycast', lambda: g.raycast(rays)), ('overlap', lambda: g.capsuleOverlap(caps)), ('overlap_all', lambda: g.capsuleOverlapAll(caps)), ('cast', lambda: g.capsuleCast(casts))):
    fn(); t=time.perf_counter(); fn(); dt=time.perf_counter()-t
    print(name, 'e2e M/s', n/dt/1e6, flush=True)
