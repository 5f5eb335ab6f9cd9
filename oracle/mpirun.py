"""oracle/mpirun.py -- TEST INFRASTRUCTURE ONLY.

`mpirun -n N` substitute for the shared-memory MPI stub (oracle/stubs/mpi_stub.c): N processes on this host, each with
CLB_MPI_RANK / CLB_MPI_NTASKS / CLB_MPI_SHM set before the reference library makes its first MPI call.  The worker is
a picklable module-level function ``fn(rank, ntasks, *args)``; results come back in rank order.
"""
import ctypes as C
import multiprocessing as mp
import os
import tempfile
import traceback


def _worker(rank, ntasks, shm, fn, args, q, extra_env):
    os.environ["CLB_MPI_RANK"] = str(rank)
    os.environ["CLB_MPI_NTASKS"] = str(ntasks)
    os.environ["CLB_MPI_SHM"] = shm
    os.environ.update(extra_env or {})
    try:
        q.put((rank, True, fn(rank, ntasks, *args)))
    except BaseException:
        q.put((rank, False, traceback.format_exc()))


def shm_bytes(ntasks):
    from . import ref
    L = ref.lib()
    L.clb_mpi_shm_bytes.restype = C.c_size_t; L.clb_mpi_shm_bytes.argtypes = [C.c_int]
    return int(L.clb_mpi_shm_bytes(ntasks))


def run(ntasks, fn, *args, timeout=1800, extra_env=None, start_method="spawn"):
    """Run fn(rank, ntasks, *args) on ntasks processes; returns the list of results by rank.  'spawn' keeps the
    children free of the parent's CUDA / library state."""
    nbytes = shm_bytes(ntasks)
    fd, shm = tempfile.mkstemp(prefix="clb_mpi_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        os.ftruncate(fd, nbytes)     # sparse, zero filled
        os.close(fd)
        ctx = mp.get_context(start_method)
        q = ctx.Queue()
        procs = [ctx.Process(target=_worker, args=(r, ntasks, shm, fn, args, q, extra_env)) for r in range(ntasks)]
        for p in procs:
            p.start()
        out = [None] * ntasks
        err = None
        for _ in range(ntasks):
            try:
                rank, ok, val = q.get(timeout=timeout)
            except Exception:
                err = "timeout waiting for the ranks"
                break
            if ok:
                out[rank] = val
            else:
                err = "rank %d failed:\n%s" % (rank, val)
                break
        for p in procs:
            if err:
                p.terminate()
            p.join(timeout=30)
        if err:
            raise RuntimeError(err)
        return out
    finally:
        try:
            os.unlink(shm)
        except OSError:
            pass
