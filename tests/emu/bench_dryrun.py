"""TEST INFRASTRUCTURE ONLY. Dry run of bench.py's GPU arm on the CPU: the library is the SIMT emulation build and
torch.cuda streams / events / pinned memory are stubbed, so the control flow, buffer handling and JSON assembly of
the bench can be exercised before the GPU box sees it. Timings printed by a dry run are meaningless."""
import contextlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ["B2P_EMU_TESTS"] = "1"


def main():
    import torch

    from tests.emu import emu_mode

    emu_mode.enable()

    class Ev:
        def __init__(self, enable_timing=False):
            pass

        def record(self, stream=None):
            pass

        def elapsed_time(self, other):
            return 1.0

    class St:
        cuda_stream = 0

        def wait_event(self, ev):
            pass

    torch.cuda.Event = Ev
    torch.cuda.Stream = St
    torch.cuda.current_stream = lambda *a, **k: St()
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.cuda.set_device = lambda *a, **k: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    _copy = torch.Tensor.copy_
    torch.Tensor.copy_ = lambda self, src, non_blocking=False: _copy(self, src)
    _zeros, _empty = torch.zeros, torch.empty
    import bench

    bench.ClockSampler = lambda *a, **k: type("C", (), {"stop": lambda self, a, b: {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["dry run"]}})()
    if len(sys.argv) > 1 and sys.argv[1].endswith(".py"):
        # dry run of another GPU tool:  python tests/emu/bench_dryrun.py tools/zfused_bench.py --n 3 ...
        import runpy

        script = sys.argv[1]
        sys.argv = [script] + sys.argv[2:]
        runpy.run_path(script, run_name="__main__")
        return
    os.environ.setdefault("B2P_BENCH_CHILD", os.path.abspath(__file__))   # experiments re-enter this harness
    if "--no-experiments" in sys.argv:                                     # (child of an experiment: bench args come from the parent)
        sys.argv = ["bench.py"] + sys.argv[1:]
    else:
        sys.argv = ["bench.py", "--n", "3", "--steps", "3", "--warmup", "1", "--cpu-sample-elems", "27"] + sys.argv[1:]
    bench.main()


if __name__ == "__main__":
    main()
