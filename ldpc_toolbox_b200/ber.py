"""BER test driver on top of the on-device BER engine (C-ABI ldpc_toolbox_ber_*).

Host-side mirror of the reference's simulation layer:
  BerTestParameters / BerTest.new      reference src/simulation/ber.rs:96-158, :246-282
  BerTest.run / do_run (stop rule)     reference src/simulation/ber.rs:288-368, :522-531
  Statistics / CodeStatistics          reference src/simulation/ber.rs:160-195, :551-581
  Modulation names                     reference src/simulation/factory.rs:56-86 (BPSK, 8PSK)

The per-frame work (random message, encode, puncture, BPSK, AWGN, decode, error counting) runs on
the GPU; this module only shards global frame indices over GPUs / ranks, sums nine counters and
applies the reference's stop rule at batch granularity.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import threading
import time
from typing import Callable, Optional, Sequence

import numpy as np

from . import capi

COUNTER_NAMES = ("frames", "bit_errors", "frame_errors", "false_decodes", "total_iterations", "correct_iterations",
                 "bch_bit_errors", "bch_frame_errors", "bch_correct_iterations")
NUM_COUNTERS = len(COUNTER_NAMES)


class BerEngine:
    """One GPU's engine: simulate a range of global frame indices at one Eb/N0, add to counters."""

    def __init__(self, alist: str, implementation: str = "Phif64", puncturing: str = "", device: int = -1, max_tiles: int = 0,
                 modulation: str = "BPSK", interleaving: Optional[int] = None):
        import os
        self._lib = capi.load()
        is_path = int("\n" not in alist and os.path.exists(alist))
        self._h = self._lib.ldpc_toolbox_ber_ctor(alist.encode(), is_path, implementation.encode(), puncturing.encode(), device, max_tiles)
        if not self._h:
            raise ValueError(f"ldpc_toolbox_ber_ctor returned NULL: {capi.last_error()}")
        if modulation != "BPSK" or interleaving:
            if self._lib.ldpc_toolbox_ber_set_modulation(self._h, modulation.encode(), int(interleaving or 0)) != 0:
                err = capi.last_error()
                self.close()
                raise ValueError(f"ldpc_toolbox_ber_set_modulation: {err}")
        dims = (C.c_uint64 * 3)()
        self._lib.ldpc_toolbox_ber_dims(self._h, dims)
        self.k, self.n_cw, self.n = int(dims[0]), int(dims[1]), int(dims[2])
        self.rate = self._lib.ldpc_toolbox_ber_rate(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ldpc_toolbox_ber_dtor(self._h)
            self._h = None

    __del__ = close

    def noise_sigma(self, ebn0_db: float) -> float:
        return self._lib.ldpc_toolbox_ber_noise_sigma(self._h, ebn0_db)

    def run(self, ebn0_db: float, max_iterations: int, first_frame: int, nframes: int, seed: int = 0x5EED,
            bch_max_errors: int = 0, counters: Optional[np.ndarray] = None) -> np.ndarray:
        if counters is None:
            counters = np.zeros(NUM_COUNTERS, dtype=np.uint64)
        rc = self._lib.ldpc_toolbox_ber_run(self._h, ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors, counters.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ldpc_toolbox_ber_run: {capi.last_error()}")
        return counters

    def submit(self, ebn0_db: float, max_iterations: int, first_frame: int, nframes: int, seed: int = 0x5EED, bch_max_errors: int = 0) -> int:
        """Asynchronous run(): enqueue the batch and return a ticket; at most two may be in flight."""
        t = self._lib.ldpc_toolbox_ber_submit(self._h, ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors)
        if t < 0:
            raise RuntimeError(f"ldpc_toolbox_ber_submit: {capi.last_error()}")
        return int(t)

    def wait(self, ticket: int, counters: Optional[np.ndarray] = None) -> np.ndarray:
        """Block until `ticket` is done and add its nine counters."""
        if counters is None:
            counters = np.zeros(NUM_COUNTERS, dtype=np.uint64)
        if self._lib.ldpc_toolbox_ber_wait(self._h, ticket, counters.ctypes.data) != 0:
            raise RuntimeError(f"ldpc_toolbox_ber_wait: {capi.last_error()}")
        return counters

    def run_dump(self, ebn0_db: float, max_iterations: int, first_frame: int, nframes: int, seed: int = 0x5EED, bch_max_errors: int = 0):
        """Test hook: also returns the LLRs, decoded info bits, iteration counts and messages."""
        counters = np.zeros(NUM_COUNTERS, dtype=np.uint64)
        llrs = np.zeros((nframes, self.n), dtype=np.float32)
        dec = np.zeros((nframes, self.k), dtype=np.uint8)
        its = np.zeros(nframes, dtype=np.int32)
        kw = (self.k + 31) // 32
        msg = np.zeros((nframes, kw), dtype=np.uint32)
        rc = self._lib.ldpc_toolbox_ber_run_dump(self._h, ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors,
                                                 counters.ctypes.data, llrs.ctypes.data, dec.ctypes.data, its.ctypes.data, msg.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"ldpc_toolbox_ber_run_dump: {capi.last_error()}")
        bits = np.unpackbits(msg.view(np.uint8).reshape(nframes, -1), axis=1, bitorder="little")[:, :self.k]
        return dict(zip(COUNTER_NAMES, counters.tolist())), llrs, dec, its, bits


# ------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class CodeStatistics:            # reference ber.rs:181-195, :551-563
    bit_errors: int
    frame_errors: int
    correct_iterations: int
    ber: float
    fer: float
    average_iterations_correct: float


@dataclasses.dataclass
class Statistics:                # reference ber.rs:160-179, :565-581
    ebn0_db: float
    num_frames: int
    false_decodes: int
    total_iterations: int
    average_iterations: float
    elapsed: float
    throughput_mbps: float
    ldpc: CodeStatistics
    bch: Optional[CodeStatistics]

    @staticmethod
    def from_counters(c: Sequence[int], ebn0_db: float, k: int, elapsed: float, has_bch: bool) -> "Statistics":
        frames = int(c[0])

        def code(bit_errors, frame_errors, correct_iterations):
            good = frames - frame_errors
            return CodeStatistics(bit_errors, frame_errors, correct_iterations,
                                  bit_errors / (k * frames) if frames else float("nan"),
                                  frame_errors / frames if frames else float("nan"),
                                  correct_iterations / good if good else float("nan"))
        return Statistics(ebn0_db, frames, int(c[3]), int(c[4]), int(c[4]) / frames if frames else float("nan"), elapsed,
                          1e-6 * k * frames / elapsed if elapsed > 0 else float("nan"),
                          code(int(c[1]), int(c[2]), int(c[5])),
                          code(int(c[6]), int(c[7]), int(c[8])) if has_bch else None)


def frame_range(launch: int, rank: int, world: int, batch: int):
    """Global frame indices of (launch, rank): f = launch*(world*batch) + rank*batch + lane  (SURVEY.md §8e)."""
    first = (launch * world + rank) * batch
    return first, batch


def run_finished(errors: int, max_frame_errors: int, elapsed: float, min_time: float, max_time: float) -> bool:
    """reference ber.rs:522-531"""
    return (errors >= max_frame_errors and elapsed >= min_time) or elapsed >= max_time


class BerTest:
    """Eb/N0 sweep (reference BerTest::do_run, ber.rs:297-368) over one or more engines.

    `engines` are objects with .run(ebn0_db, max_iterations, first_frame, nframes, seed, bch_max_errors, counters);
    one per local GPU.  `allreduce(counters)` (optional) sums the nine counters over ranks of a
    torch.distributed job so every rank applies the stop rule to the same global counts.
    """

    def __init__(self, engines, k: int, ebn0s_db: Sequence[float], max_iterations: int = 100, max_frame_errors: int = 100,
                 min_time: float = 0.0, max_time: float = float("inf"), bch_max_errors: int = 0, batch: int = 4096, seed: int = 0x5EED,
                 rank: int = 0, world: int = 1, allreduce: Optional[Callable[[np.ndarray], np.ndarray]] = None,
                 reporter: Optional[Callable[[Statistics, bool], None]] = None, max_frames: Optional[int] = None):
        self.engines = list(engines)
        self.k = k
        self.ebn0s_db = [float(np.float32(e)) for e in ebn0s_db]
        self.max_iterations, self.max_frame_errors = max_iterations, max_frame_errors
        self.min_time, self.max_time = min_time, max_time
        self.bch_max_errors = bch_max_errors
        self.batch, self.seed = batch, seed
        self.rank, self.world, self.allreduce = rank, world, allreduce
        self.reporter = reporter
        self.max_frames = max_frames
        self.statistics: list[Statistics] = []

    @property
    def pipeline_depth(self) -> int:
        """Rounds kept in flight: 2 when every engine has the asynchronous submit/wait pair, else 1."""
        return 2 if all(hasattr(e, "submit") and hasattr(e, "wait") for e in self.engines) else 1

    def overshoot_bound(self) -> int:
        """Frames that may be simulated beyond the round in which the stop rule first holds."""
        return (self.pipeline_depth - 1) * self.batch * len(self.engines) * self.world

    def _submit_round(self, ebn0_db: float, launch: int):
        """Enqueue one round on every local engine (asynchronous engines) or run it (blocking ones)."""
        nloc = len(self.engines)
        if self.pipeline_depth == 2:
            tickets = []
            for i, eng in enumerate(self.engines):
                first, n = frame_range(launch, self.rank * nloc + i, self.world * nloc, self.batch)
                tickets.append(eng.submit(ebn0_db, self.max_iterations, first, n, self.seed, self.bch_max_errors))
            return ("tickets", tickets)
        parts = [np.zeros(NUM_COUNTERS, dtype=np.uint64) for _ in range(nloc)]
        errs: list = [None] * nloc

        def work(i):
            try:
                first, n = frame_range(launch, self.rank * nloc + i, self.world * nloc, self.batch)
                self.engines[i].run(ebn0_db, self.max_iterations, first, n, self.seed, self.bch_max_errors, parts[i])
            except Exception as e:   # a worker error aborts the sweep (ber.rs:339,:351-359)
                errs[i] = e
        if nloc == 1:
            work(0)
        else:
            th = [threading.Thread(target=work, args=(i,)) for i in range(nloc)]
            [t.start() for t in th]
            [t.join() for t in th]
        for e in errs:
            if e is not None:
                raise e
        return ("counters", parts)

    def _collect_round(self, pending) -> np.ndarray:
        """Wait for a submitted round, sum its counters over the local engines and (multi-rank) over ranks.  With two
        rounds in flight this host work — and the all-reduce — overlaps the kernels of the next round."""
        kind, items = pending
        if kind == "tickets":
            parts = [eng.wait(t) for eng, t in zip(self.engines, items)]
        else:
            parts = items
        total = np.sum(parts, axis=0).astype(np.uint64)
        if self.allreduce is not None:
            # The stop decision must be collective: besides the nine counters, rank 0's clock travels in
            # the same all-reduce, so every rank applies ber.rs:522-531 to identical numbers and leaves
            # the loop on the same round (the reference has one clock, in the controlling thread).
            ext = np.zeros(NUM_COUNTERS + 1, dtype=np.uint64)
            ext[:NUM_COUNTERS] = total
            if self.rank == 0:
                ext[NUM_COUNTERS] = np.uint64(max(0.0, time.perf_counter() - self._start) * 1e6)
            ext = self.allreduce(ext)
            total = ext[:NUM_COUNTERS].astype(np.uint64)
            self._shared_elapsed = float(ext[NUM_COUNTERS]) * 1e-6
        return total

    def run(self) -> list[Statistics]:
        has_bch = self.bch_max_errors > 0
        depth = self.pipeline_depth
        frames_per_round = self.batch * len(self.engines) * self.world
        for ebn0_db in self.ebn0s_db:
            counters = np.zeros(NUM_COUNTERS, dtype=np.uint64)
            start = self._start = time.perf_counter()
            self._shared_elapsed = 0.0
            launch, submitted = 0, 0
            pending: list = []
            stopping = False
            while True:
                # multi-rank: rank 0's clock as of the last all-reduce; single process: the local clock
                elapsed = self._shared_elapsed if self.allreduce is not None else time.perf_counter() - start
                errors = int(counters[7] if has_bch else counters[2])        # ber.rs:514-520
                if run_finished(errors, self.max_frame_errors, elapsed, self.min_time, self.max_time):
                    stopping = True
                if self.max_frames is not None and int(counters[0]) >= self.max_frames:
                    stopping = True
                # keep `depth` rounds in flight; once the stop rule holds, rounds already submitted are still collected
                while not stopping and len(pending) < depth and (self.max_frames is None or submitted < self.max_frames):
                    pending.append(self._submit_round(ebn0_db, launch))
                    launch += 1
                    submitted += frames_per_round
                if not pending:
                    break
                counters += self._collect_round(pending.pop(0))
                if self.reporter:
                    self.reporter(Statistics.from_counters(counters, ebn0_db, self.k, time.perf_counter() - start, has_bch), False)
            st = Statistics.from_counters(counters, ebn0_db, self.k, time.perf_counter() - start, has_bch)
            if self.reporter:
                self.reporter(st, True)
            self.statistics.append(st)
        return self.statistics
