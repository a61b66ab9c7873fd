import importlib, sys, time, os
sys.path.insert(0,'.')
cq = importlib.import_module('swift-game-engine_b200')
parts, half = cq.scenes.terrain_scene(cells=2236, cell=2.0)
for mode in ('onesweep','classic','onesweep','classic'):
    os.environ['CQ_SORT']=mode
    for k in range(2):
        t=time.perf_counter(); g = cq.CollisionQuery(parts); dt=time.perf_counter()-t
        print(mode, k, 'create_s %.3f'%dt, 'build_ms %.1f'%g.info()['build_ms'], flush=True)
        g.close()
