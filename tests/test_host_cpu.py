"""CPU: host-side logic added in round 2 that needs no GPU -- the prefetching loader's pass-through / ordering,
compact host batches, the deferred validity checks of the training loop."""
import pytest
import torch

from sgs_gnn_b200 import _train_core, loader, synth


def test_prefetch_passes_cpu_and_resident_batches_through_in_order():
    bs = [synth.make_graph(None, seed=s, n=40, e=160, f=4, c=2) for s in range(3)]
    out = list(loader.prefetch(bs, "cpu"))
    assert len(out) == 3 and all(a is b for a, b in zip(out, bs))
    assert list(loader.prefetch([], "cpu")) == []
    # a generator loader is consumed lazily, one element ahead
    seen = []

    def gen():
        for b in bs:
            seen.append(1)
            yield b
    it = loader.prefetch(gen(), "cpu")
    first = next(it)
    assert first is bs[0] and len(seen) == 2
    assert [b is c for b, c in zip(list(it), bs[1:])] == [True, True]


def test_compact_batch_keeps_everything_but_the_index_dtype():
    b = synth.make_graph(None, seed=1, n=50, e=200, f=4, c=2)
    c = b.compact()
    assert c.edge_index.dtype == torch.int32 and torch.equal(c.edge_index.long(), b.edge_index)
    assert c.x is b.x and c.prob is b.prob and c.num_classes == b.num_classes
    assert c.nbytes() == b.nbytes() - b.edge_index.numel() * 4
    assert c.compact().edge_index.dtype == torch.int32


def test_deferred_checks_raise_like_the_reference():
    loss = torch.tensor(1.25)
    ok = torch.tensor([0, 0, 0, 0, 0, 0, 0, 7], dtype=torch.int64)
    assert _train_core._read_loss_and_checks(loss, [], 7) == 1.25
    assert _train_core._read_loss_and_checks(loss, [("sampler", ok), ("oob", torch.zeros(1, dtype=torch.int32))], 7) == 1.25
    bad = ok.clone()
    bad[5] = 1
    with pytest.raises(RuntimeError, match="inf"):
        _train_core._read_loss_and_checks(loss, [("sampler", bad)], 7)
    with pytest.raises(RuntimeError, match="expected 7"):
        _train_core._read_loss_and_checks(loss, [("sampler", torch.tensor([0, 0, 0, 0, 0, 0, 0, 6]))], 7)
    with pytest.raises(RuntimeError, match="outside"):
        _train_core._read_loss_and_checks(loss, [("oob", torch.ones(1, dtype=torch.int32))], 7)


def test_bench_clock_sampler_keeps_the_samples_of_the_timed_region(tmp_path):
    """bench.py's ClockSampler: samples carry nvidia-smi timestamps; only those inside [begin, end] count, and a
    region too short to hold one falls back to the warm-up steps just before it (and says so)."""
    import datetime
    import importlib.util
    import os
    import time
    spec = importlib.util.spec_from_file_location("sgs_bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    now = time.time()

    def write(path):
        with open(path, "w") as fh:
            for i in range(10):
                ts = datetime.datetime.fromtimestamp(now - 1 + 0.1 * i).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
                cap = "Active" if i == 7 else "Not Active"
                fh.write(f"{ts}, {1900 + i}, 1965, 800.1, Not Active, Not Active, Not Active, {cap}\n")

    c = bench.ClockSampler(0)
    c.path = str(tmp_path / "a.csv")
    write(c.path)
    c.t0, c.t1 = now - 0.45, now - 0.28
    out = c.summary()
    assert out["window"] == "timed" and out["samples"] == 2 and out["sm_mhz"] == 1906.5
    assert out["reasons"] == ["sw_power_cap"] and out["sm_max_mhz"] == 1965.0
    c = bench.ClockSampler(0)
    c.path = str(tmp_path / "b.csv")
    write(c.path)
    c.t0, c.t1 = now + 0.5, now + 0.52           # nothing inside: the 3 s before it
    out = c.summary()
    assert out["window"] == "warmup+timed" and out["samples"] == 10


def test_bench_clock_sampler_reads_a_live_child_line_by_line():
    """The child writes to a pseudo-terminal, so every sample arrives as soon as it is printed (a file or pipe would
    hold ~4 KB back and lose it at terminate)."""
    import importlib.util
    import os
    import time
    spec = importlib.util.spec_from_file_location("sgs_bench", os.path.join(os.path.dirname(__file__), "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = bench.ClockSampler(0)
    c.cmd = ["sh", "-c", 'while true; do echo "$(date "+%Y/%m/%d %H:%M:%S.%3N"), 1900, 1965, 800.0, Not Active, '
             'Not Active, Not Active, Active"; sleep 0.05; done']
    with c:
        assert c.wait_ready(5.0)
        c.begin()
        time.sleep(0.4)
        c.end()
    out = c.summary()
    assert out["window"] == "timed" and out["samples"] >= 3 and out["sm_mhz"] == 1900.0
    assert out["reasons"] == ["sw_power_cap"]
    assert c.proc.poll() is not None        # the child is gone


def test_bind_host_to_gpu_is_a_hint_not_an_error():
    """Without a GPU / NVML the placement helper reports None and leaves the affinity alone."""
    import os
    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert loader.bind_host_to_gpu("cuda:0") is None
    assert os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)


def test_compact_sends_a_sorted_source_row_as_its_row_pointer():
    b = synth.make_graph(None, seed=3, n=60, e=400, f=4, c=2)
    c = b.compact()
    rp = c._src_rowptr
    assert rp.dtype == torch.int32 and rp.numel() == 61 and int(rp[0]) == 0 and int(rp[-1]) == 400
    deg = (rp[1:] - rp[:-1]).to(torch.int64)
    src = torch.repeat_interleave(torch.arange(60, dtype=torch.int32), deg, output_size=400)
    assert torch.equal(src, c.edge_index[0])
    assert c.upload_nbytes() == c.nbytes() - 400 * 4 + 61 * 4
    assert b.compact(rowptr=False).upload_nbytes() == c.nbytes()
    b.edge_index = b.edge_index.flip(1)            # not sorted by source: the plain int32 form
    assert getattr(b.compact(), "_src_rowptr", None) is None
