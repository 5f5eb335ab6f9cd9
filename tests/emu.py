"""Ranks emulated as threads of one process (tests): an allgather for clb_solver_create plus a runner."""
import threading


class ThreadRanks:
    def __init__(self, n, timeout=300):
        self.n = n
        self.barrier = threading.Barrier(n, timeout=timeout)
        self.slots = [None] * n

    def gather_fn(self, rank):
        def gather(data):
            self.slots[rank] = data
            self.barrier.wait()
            out = list(self.slots)
            self.barrier.wait()
            return out
        return gather

    def run(self, fn):
        """fn(rank, gather) on n threads; returns the results by rank, re-raises the first failure"""
        res, errs = [None] * self.n, []

        def body(r):
            try:
                res[r] = fn(r, self.gather_fn(r))
            except BaseException as e:   # noqa: BLE001
                errs.append(e)
                self.barrier.abort()
        th = [threading.Thread(target=body, args=(r,)) for r in range(self.n)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return res
