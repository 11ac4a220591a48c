"""Times one config-2 train step with the reference's default generator (ContextNetwork) at B=64 and
prints the per-family breakdown (developer tool; numbers quoted in DESIGN.md section 4.3)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcnet_b200 import ops                               # noqa: E402
from dmcnet_b200.engine import DmcEngine                   # noqa: E402
from dmcnet_b200.model import build_state                  # noqa: E402
from dmcnet_b200.profiling import FamilyTimer              # noqa: E402
from dmcnet_b200.trainer import FusedTrainStep, HParams    # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    att = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    sd = build_state(51, None, seed=1, arch_estimator='ContextNetwork') if not att else None
    eng = DmcEngine(51, 3, B * 3, arch_estimator='ContextNetwork', att=att)
    if sd is not None:
        eng.load_state(sd)
    else:
        torch.manual_seed(0)
        eng.params.normal_(0, 0.05)
    tr = FusedTrainStep(eng, HParams(), B, use_graph=True)
    g = torch.Generator().manual_seed(0)
    flow = torch.randn(B, 3, 2, 224, 224, generator=g).cuda()
    mv = torch.randn(B, 3, 2, 224, 224, generator=g).cuda()
    res = torch.randn(B, 3, 3, 224, 224, generator=g).cuda()
    t = torch.randint(0, 51, (B,), generator=g).cuda()
    for _ in range(4):
        tr.step(flow, mv, res, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        tr.step(flow, mv, res, t, metrics=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('ContextNetwork%s config-2 step B=%d: %.2f ms -> %.1f clips/s; peak memory %.1f GB'
          % ('Att' if att else '', B, ms, B / ms * 1e3, torch.cuda.max_memory_allocated() / 1e9))
    tr.use_graph = False
    timer = FamilyTimer()
    ops.set_call_hook(timer)
    tr.step(flow, mv, res, t, metrics=False)
    ops.set_call_hook(None)
    ms_, fl, by, cnt = timer.totals()
    for k, v in sorted(ms_.items(), key=lambda kv: -kv[1])[:12]:
        print('%-28s %8.3f ms  %7.1f TFLOP/s (issued columns)  x%d' % (k, v, fl[k] / v / 1e9 if v else 0, cnt[k]))


if __name__ == '__main__':
    main()
