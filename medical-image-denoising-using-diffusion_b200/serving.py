"""Cross-request micro-batching in front of the hot-path models.

`/denoise` as deployed calls every model with ONE 512x512 image from its own worker thread
(`asyncio.to_thread`, RUN:85-91; `_process_diffusion/_process_nafnet/_process_hybrid`, RUN:104-141).  One image
under-fills a B200: the 9-evaluation sampler takes 18 ms for a single image and 10.8 ms per evaluation *for sixteen*.  A
`MicroBatcher` sits between those worker threads and a model: requests that arrive within `max_delay_ms` of each
other (and share shape, dtype and device) are concatenated, run as one batch through the same library call, and
split again.  Images never interact inside the networks, so a request's result does not
depend on what it was batched with beyond the summation order of the GroupNorm atomics.

    batcher = MicroBatcher(lambda x: wrapper.denoise(x, 8), max_batch=16, max_delay_ms=2.0)
    out = batcher(input_tensor)            # from any thread; blocks that thread only; same result as the direct call

The callable runs on one dedicated thread (its own CUDA stream ordering is preserved with events: inputs are
consumed after the submitting stream produced them, outputs are visible to the submitting stream on return).
"""
from __future__ import annotations

import collections
import queue
import threading
import time
from concurrent.futures import Future
from typing import Callable, List, Optional, Tuple

import torch


class MicroBatcher:
    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor], max_batch: int = 16, max_delay_ms: float = 2.0):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.fn = fn
        self.max_batch = int(max_batch)
        self.max_delay = float(max_delay_ms) / 1e3
        # sizes of the most recent batches actually run (observability / tests); bounded: a server lives for months
        self.batches: "collections.deque[int]" = collections.deque(maxlen=4096)
        self._submit_lock = threading.Lock()   # orders submit() against close(): nothing is enqueued behind the stop sentinel
        self._q: "queue.Queue[Optional[Tuple[torch.Tensor, Optional[torch.cuda.Event], Future]]]" = queue.Queue()
        self._closed = False
        self._worker = threading.Thread(target=self._run, name="xrd-microbatcher", daemon=True)
        self._worker.start()

    # ------------------------------------------------------------------ client side
    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        return self.submit(x).result()

    def submit(self, x: torch.Tensor) -> Future:
        """Enqueue a (k,1,H,W) request (k >= 1); the Future resolves to the (k,1,H,W) result."""
        if not isinstance(x, torch.Tensor) or x.dim() != 4:
            raise ValueError(f"expected a (k,C,H,W) tensor, got {type(x).__name__} {tuple(getattr(x, 'shape', ()))}")
        ev = None
        if x.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(x.device))
        fut: Future = Future()
        with self._submit_lock:
            if self._closed:
                raise RuntimeError("MicroBatcher is closed")
            self._q.put((x, ev, fut))
        return fut

    def close(self) -> None:
        """Stop accepting requests, run what was already submitted, stop the worker.  Idempotent."""
        with self._submit_lock:
            if self._closed:
                return
            self._closed = True
            self._q.put(None)
        self._worker.join(timeout=30)
        # a worker that died or is stuck in a long call must not leave waiters blocked for ever
        while True:
            try:
                it = self._q.get_nowait()
            except queue.Empty:
                break
            if it is not None and not it[2].done():
                it[2].set_exception(RuntimeError("MicroBatcher closed before the request ran"))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ worker side
    @staticmethod
    def _key(x: torch.Tensor):
        return (tuple(x.shape[1:]), x.dtype, x.device)

    def _run(self) -> None:
        pending: List[Tuple[torch.Tensor, Optional[torch.cuda.Event], Future]] = []   # requests set aside for the next batch
        while True:
            item = pending.pop(0) if pending else self._q.get()
            if item is None:
                break
            group = [item]
            n = item[0].shape[0]
            key = self._key(item[0])
            deadline = time.monotonic() + self.max_delay
            stop = False
            # requests set aside earlier that match this group join it first, in arrival order
            rest = []
            for it in pending:
                if self._key(it[0]) == key and n + it[0].shape[0] <= self.max_batch:
                    group.append(it); n += it[0].shape[0]
                else:
                    rest.append(it)
            pending = rest
            while n < self.max_batch:
                wait = deadline - time.monotonic()
                if wait <= 0:
                    break
                try:
                    nxt = self._q.get(timeout=wait)
                except queue.Empty:
                    break
                if nxt is None:
                    stop = True
                    break
                if self._key(nxt[0]) == key and n + nxt[0].shape[0] <= self.max_batch:
                    group.append(nxt); n += nxt[0].shape[0]
                else:
                    pending.append(nxt)            # another shape (or overflow): it opens the next batch
            self._run_group(group)
            if stop:
                for it in pending:
                    self._run_group([it])
                break

    def _run_group(self, group) -> None:
        xs = [g[0] for g in group]
        try:
            dev = xs[0].device
            if dev.type == "cuda":
                stream = torch.cuda.current_stream(dev)
                for _, ev, _ in group:
                    if ev is not None:
                        stream.wait_event(ev)
            batch = xs[0] if len(xs) == 1 else torch.cat(xs, 0)
            out = self.fn(batch)
            self.batches.append(int(batch.shape[0]))
            if dev.type == "cuda":
                torch.cuda.current_stream(dev).synchronize()     # results are complete before any waiter wakes up
            off = 0
            for x, _, fut in group:
                k = x.shape[0]
                fut.set_result(out[off:off + k])
                off += k
        except BaseException as e:  # noqa: BLE001 - every waiter must be released, with the error (RUN:96-101 turns it into a null result)
            for _, _, fut in group:
                if not fut.done():
                    fut.set_exception(e)
