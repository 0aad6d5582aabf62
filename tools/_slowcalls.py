import sys, time, io, contextlib, warnings
import numpy as np, torch
sys.path.insert(0, '/root/repo')
warnings.filterwarnings("ignore")
import hmvec_b200 as hm
zs = np.linspace(0.01, 3., 200); ms = np.geomspace(2e10, 1e17, 2000); ks = np.geomspace(1e-4, 100, 10000)
ngal = np.geomspace(1e-3, 1e-5, 200)
log = []
def wrap(mod, name):
    f = getattr(mod, name)
    def g(*a, **k):
        t0 = time.perf_counter(); r = f(*a, **k); dt = time.perf_counter() - t0
        if dt > 3e-4: log.append((name, round(dt * 1e3, 2), str(a[:2])[:60], str(k)[:60]))
        return r
    setattr(mod, name, g)
wrap(torch, 'empty'); wrap(torch, 'zeros'); wrap(torch.Tensor, 'to'); wrap(torch.Tensor, 'copy_'); wrap(torch.Tensor, 'record_stream')
wrap(torch.cuda.Stream, 'wait_event'); wrap(torch.cuda.Event, 'record'); wrap(torch.cuda.Stream, '__new__')
def step():
    with contextlib.redirect_stdout(io.StringIO()):
        h = hm.HaloModel(zs, ks, ms=ms, accuracy='low')
        h.add_battaglia_profile("electron", family="AGN", xmax=20, nxs=5000)
        h.add_battaglia_pres_profile("y", family="pres", xmax=20, nxs=5000)
        t0 = time.perf_counter(); log.append(("-- add_hod begins", 0, "", ""))
        h.add_hod("g", ngal=ngal)
        log.append(("-- add_hod ends", round((time.perf_counter() - t0) * 1e3, 2), "", ""))
        P = h.get_power("nfw", "nfw")
    torch.cuda.synchronize()
for i in range(4):
    log.clear(); step()
for l in log: print(l)
