import sys, time, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests/golden')
import torch
from bench import cpu_step_factory
for th in (16, 32, 64, 128):
    step, st = cpu_step_factory(20, 416, 4, th)
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); step(); t2 = time.perf_counter()
    print('threads', th, 'step1 %.1fs step2 %.1fs' % (t1 - t0, t2 - t1), flush=True)
