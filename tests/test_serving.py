"""Cross-request micro-batching: host logic on CPU with a stand-in model, and on the GPU the
hybrid called from concurrent threads, as run.py's asyncio.to_thread workers do (RUN:85-91)."""
import threading
import time

import pytest
import torch

import xrd_b200


def _fake_model(calls):
    def fn(x):
        calls.append(int(x.shape[0]))
        time.sleep(0.01)
        return x * 2.0 + x.mean(dim=(1, 2, 3), keepdim=True)     # per-image: independent of the batch composition
    return fn


def test_requests_from_threads_are_merged_and_answered_in_order():
    calls = []
    fn = _fake_model(calls)
    xs = [torch.full((1, 1, 8, 8), float(i)) + torch.arange(64.0).view(1, 1, 8, 8) for i in range(12)]
    outs = [None] * len(xs)
    with xrd_b200.MicroBatcher(fn, max_batch=8, max_delay_ms=200.0) as mb:
        def work(i):
            outs[i] = mb(xs[i])
        ts = [threading.Thread(target=work, args=(i,)) for i in range(len(xs))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    assert sum(calls) == 12 and max(calls) <= 8 and len(calls) < 12          # merged, never above max_batch
    for i, x in enumerate(xs):
        assert torch.equal(outs[i], fn(x))                                   # each caller gets its own image back


def test_shapes_are_not_mixed_and_multi_image_requests_stay_whole():
    calls = []
    with xrd_b200.MicroBatcher(_fake_model(calls), max_batch=4, max_delay_ms=100.0) as mb:
        a = mb.submit(torch.ones(2, 1, 8, 8))
        b = mb.submit(torch.ones(1, 1, 16, 16))
        c = mb.submit(torch.ones(3, 1, 8, 8))          # 2 + 3 > max_batch: opens its own batch
        d = mb.submit(torch.ones(1, 1, 8, 8))
        assert a.result().shape == (2, 1, 8, 8) and b.result().shape == (1, 1, 16, 16)
        assert c.result().shape == (3, 1, 8, 8) and d.result().shape == (1, 1, 8, 8)
    assert sum(calls) == 7 and max(calls) <= 4


def test_close_runs_what_was_submitted_and_refuses_later_requests():
    calls = []
    mb = xrd_b200.MicroBatcher(_fake_model(calls), max_batch=4, max_delay_ms=500.0)
    futs = [mb.submit(torch.ones(1, 1, 8, 8) * i) for i in range(3)]
    mb.close()                                          # the stop sentinel sits behind the three requests
    assert all(f.done() for f in futs) and sum(calls) == 3
    with pytest.raises(RuntimeError):
        mb.submit(torch.ones(1, 1, 8, 8))
    mb.close()                                          # idempotent
    assert mb.batches.maxlen is not None                # the batch-size log is bounded


def test_errors_reach_every_waiter_and_the_batcher_survives():
    def bad(x):
        if x.shape[-1] == 4:
            raise xrd_b200.XrdError("unsupported shape")
        return x
    with xrd_b200.MicroBatcher(bad, max_batch=4, max_delay_ms=50.0) as mb:
        f1, f2 = mb.submit(torch.ones(1, 1, 4, 4)), mb.submit(torch.ones(1, 1, 4, 4))
        with pytest.raises(xrd_b200.XrdError):
            f1.result()
        with pytest.raises(xrd_b200.XrdError):
            f2.result()
        assert torch.equal(mb(torch.ones(1, 1, 8, 8)), torch.ones(1, 1, 8, 8))
    with pytest.raises(RuntimeError):
        mb.submit(torch.ones(1, 1, 8, 8))               # closed
    with pytest.raises(ValueError):
        xrd_b200.MicroBatcher(bad, max_batch=0)


@pytest.mark.gpu
def test_concurrent_single_image_requests_equal_direct_calls():
    import gpu_checks as G
    from oracle import xrd_oracle as O
    m, _ = G._hybrid("fp16")
    m.inference_diffusion_steps = 8                      # the served configuration (RUN:72-73): 9 evaluations
    _, noisy = O.synthetic_xray(6, 128, 128, seed=17)
    xs = [noisy[i:i + 1].to(G.DEV) for i in range(6)]
    direct = [m(x) for x in xs]
    outs = [None] * 6
    with xrd_b200.MicroBatcher(m, max_batch=16, max_delay_ms=100.0) as mb:
        def work(i):
            outs[i] = mb(xs[i])
        ts = [threading.Thread(target=work, args=(i,)) for i in range(6)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert sum(mb.batches) == 6 and len(mb.batches) < 6
    for i in range(6):
        assert (outs[i] - direct[i]).abs().max() < 5e-3      # identical up to the GroupNorm atomics' summation order


@pytest.mark.gpu
def test_three_models_from_three_threads_like_run_py():
    """RUN:85-91: the sampler, NAFNet and the hybrid run concurrently from three OS threads on one GPU, each model object with
    its own library handle (graph capture is thread-local, handles share nothing).  Results must equal the sequential calls."""
    import gpu_checks as G
    from oracle import xrd_oracle as O
    hyb, _ = G._hybrid("fp16")
    hyb.inference_diffusion_steps = 8
    unet, _ = G.seeded_state_dict("unet")
    unet = unet.to(G.DEV)
    wrapper = xrd_b200.DiffusionDenoiser(unet, noise_steps=50)
    naf, _ = G.seeded_state_dict("nafnet")
    naf = naf.to(G.DEV)
    _, noisy = O.synthetic_xray(1, 64, 64, seed=29)
    x = noisy.to(G.DEV)
    jobs = {"diffusion": lambda: wrapper.denoise(x, inference_steps=8), "nafnet": lambda: naf(x), "hybrid": lambda: hyb(x)}
    seq = {k: f().clone() for k, f in jobs.items()}
    for _ in range(3):                                   # first round captures the graphs concurrently, later rounds replay them
        out, err = {}, {}

        def work(k):
            try:
                out[k] = jobs[k]()
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                err[k] = e
        ts = [threading.Thread(target=work, args=(k,)) for k in jobs]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        assert not err, err
        for k in jobs:
            assert (out[k] - seq[k]).abs().max() < 5e-3, k
