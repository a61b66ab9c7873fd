import importlib, sys, numpy as np
sys.path.insert(0,'.')
cq = importlib.import_module('swift-game-engine_b200')
for mesh in ('hulls','render'):
    parts = cq.scenes.mirror_scene(use_hulls=(mesh=='hulls'))
    n = 262144 if mesh=='hulls' else 65536
    pos, vel = cq.scenes.gen_c3_characters(n, seed=0xC0111DE3)
    g = cq.CollisionQuery(parts)
    s = cq.init_states(pos, vel); p = cq.default_params()
    for _ in range(3): g.move_and_slide(s, p)
    g.set_counting(True); g.resetStats()
    g.move_and_slide(s, p)
    c = g.stats()
    ev = s['_pad'][:,1].astype(np.uint32) | (s['_pad'][:,2].astype(np.uint32)<<8) | (s['_pad'][:,3].astype(np.uint32)<<16)
    print(mesh, c, 'sum', ev.sum())
    print(' per-char evals: mean %.1f p50 %d p90 %d p99 %d p99.9 %d max %d' % (ev.mean(), *np.percentile(ev,[50,90,99,99.9]), ev.max()))
    w = ev[: (n//32)*32].reshape(-1,32)
    print(' warp max/mean (static, consecutive 32):', (w.max(1).mean()/w.mean()))
    hist = np.histogram(ev, bins=[0,25,50,100,200,400,800,1600,3200,1e9])[0]
    print(' hist', hist/len(ev))
