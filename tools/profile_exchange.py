import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vpic_b200 import engine as E
class A: pass
args=A(); args.grid=128; args.ppc=64; args.uth=0.18; args.sort_interval="20"; args.variant=0
rank=int(os.environ['RANK']); world=int(os.environ['WORLD_SIZE']); local=int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local); dev=torch.device('cuda',local)
dist.init_process_group('nccl', device_id=dev)
sim=bench.build_sim(args, rank, world, dev)
for _ in range(3): sim.advance()
T={}
def timed(name, fn):
    torch.cuda.synchronize(); t0=time.perf_counter(); fn(); torch.cuda.synchronize(); T[name]=T.get(name,0)+time.perf_counter()-t0
fa, ia, aa = sim.field_array, sim.interpolator_array, sim.accumulator_array
ex=sim.exchange
N=10
for it in range(N):
    timed('clear', lambda: E.clear_accumulator_array(aa))
    def push():
        for sp in sim.species_list: E.advance_p(sp, aa, ia, sync=False)
    timed('advance_p', push)
    def fin():
        for sp in sim.species_list: E.finish_advance_p(sp)
    timed('finish(nm+sort movers)', fin)
    for r in range(3):
        timed(f'boundary_p round {r}', lambda: ex.boundary_p(sim))
    timed('clear_jf+unload+syncjf_local', lambda: (fa.clear_jf(), E.unload_accumulator_array(fa, aa), fa.synchronize_jf()))
    timed('halo jf', lambda: ex.synchronize_jf(sim))
    timed('advance_b', lambda: fa.advance_b(0.5))
    timed('halo tang_b', lambda: ex.ghost_tang_b(sim))
    timed('advance_e+b', lambda: (fa.advance_e(1.0), fa.advance_b(0.5)))
    timed('load_interp', lambda: E.load_interpolator_array(ia, fa))
    sim.g.g.step += 1
if rank==0:
    for k,v in T.items(): print(f"{k:35s} {1e3*v/N:8.3f} ms")
    print("movers:", [ (sp.np) for sp in sim.species_list])
dist.destroy_process_group()
