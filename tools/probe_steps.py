import importlib, sys, numpy as np, torch
sys.path.insert(0,'.')
cq = importlib.import_module('swift-game-engine_b200')
parts = cq.scenes.mirror_scene(use_hulls=True)
world = cq.CollisionQuery(parts)
params = cq.default_params()
dev = torch.device('cuda',0)
st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st)
for n in (131072, 262144, 524288, 1048576, 2097152):
    pos, vel = cq.scenes.gen_c3_characters(n, seed=0xC0111DE3)
    s0 = cq.init_states(pos, vel)
    d = torch.from_numpy(s0.view(np.uint8).reshape(-1).copy()).to(dev)
    ms = []
    for k in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); world.move_and_slide_device(d.data_ptr(), n, params, 1/60., (0,-98.,0), 1, st.cuda_stream); e1.record()
        torch.cuda.synchronize(); ms.append(round(e0.elapsed_time(e1),3))
    print(n, ms, 'Mchars/s', round(n/ms[-1]/1e3,1), flush=True)
