"""Stock NCCL driven through its own C API (ctypes) — the reference arm of every benchmark.

The reference ships no collective code: its nccl-test manifests mount an installer-dropped NCCL and run
nccl-tests' `*_perf` binaries (reference: gpudirect-rdma/nccl-test-a4.yaml:42-77,
gpudirect-tcpx/nccl-config.yaml:30-63). On one NVSwitch box the net plugin is never on the data path, so
"the reference's NCCL build on this box" is the image's libnccl called exactly the way nccl-tests calls it:
ncclAllReduce/ncclAllGather/ncclReduceScatter/ncclBroadcast/ncclReduce/ncclSend+ncclRecv on raw device pointers, one rank per GPU.
None of this repo's kernels are on this path.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Optional

NCCL_FLOAT16, NCCL_FLOAT32, NCCL_BFLOAT16 = 6, 7, 9
NCCL_SUM = 0

# The reference's canonical environment profile for the intra-node part (gpudirect-tcpxo/README.md:71-103).
REFERENCE_ENV = {
    "NCCL_NVLS_ENABLE": "1",
    "NCCL_PROTO": "Simple,LL128",
    "NCCL_MIN_NCHANNELS": "4",
    "NCCL_P2P_NVL_CHUNKSIZE": "1048576",
    "NCCL_BUFFSIZE": "8388608",
    "NCCL_NVLSTREE_MAX_CHUNKSIZE": "131072",
}


class NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


def find_libnccl() -> Optional[str]:
    cands = []
    try:
        import torch
        base = os.path.dirname(os.path.dirname(torch.__file__))
        cands += glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so*"))
        cands += glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libnccl.so*"))
    except Exception:
        pass
    cands += glob.glob("/usr/lib/x86_64-linux-gnu/libnccl.so*")
    return cands[0] if cands else None


class NcclComm:
    def __init__(self, rank: int, nranks: int, uid_bytes: bytes, lib_path: Optional[str] = None):
        path = lib_path or find_libnccl()
        if not path:
            raise ImportError("libnccl not found")
        self.lib = L = C.CDLL(path)
        self.path = path
        vp = C.c_void_p
        L.ncclGetErrorString.restype = C.c_char_p
        L.ncclCommInitRank.argtypes = [C.POINTER(vp), C.c_int, NcclUniqueId, C.c_int]
        L.ncclAllReduce.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, vp, vp]
        L.ncclAllGather.argtypes = [vp, vp, C.c_size_t, C.c_int, vp, vp]
        L.ncclReduceScatter.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, vp, vp]
        L.ncclBroadcast.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, vp, vp]
        L.ncclReduce.argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp, vp]
        L.ncclSend.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, vp, vp]
        L.ncclRecv.argtypes = [vp, C.c_size_t, C.c_int, C.c_int, vp, vp]
        L.ncclCommDestroy.argtypes = [vp]
        if hasattr(L, "ncclMemAlloc"):
            L.ncclMemAlloc.argtypes = [C.POINTER(vp), C.c_size_t]
            L.ncclMemFree.argtypes = [vp]
        if hasattr(L, "ncclCommWindowRegister"):
            L.ncclCommWindowRegister.argtypes = [vp, vp, C.c_size_t, C.POINTER(vp), C.c_int]
            L.ncclCommWindowDeregister.argtypes = [vp, vp]
        uid = NcclUniqueId()
        C.memmove(C.byref(uid), uid_bytes, 128)
        self.comm = vp()
        self.rank, self.nranks = rank, nranks
        self._ck(L.ncclCommInitRank(C.byref(self.comm), nranks, uid, rank), "ncclCommInitRank")
        v = C.c_int()
        L.ncclGetVersion(C.byref(v))
        self.version = v.value

    @staticmethod
    def new_unique_id(lib_path: Optional[str] = None) -> bytes:
        L = C.CDLL(lib_path or find_libnccl())
        uid = NcclUniqueId()
        rc = L.ncclGetUniqueId(C.byref(uid))
        if rc != 0:
            raise RuntimeError(f"ncclGetUniqueId -> {rc}")
        return bytes(uid)

    def _ck(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what}: {self.lib.ncclGetErrorString(rc).decode()}")

    def all_reduce(self, send: int, recv: int, count: int, dtype: int, stream: int) -> None:
        self._ck(self.lib.ncclAllReduce(send, recv, count, dtype, NCCL_SUM, self.comm, stream), "ncclAllReduce")

    def all_gather(self, send: int, recv: int, sendcount: int, dtype: int, stream: int) -> None:
        self._ck(self.lib.ncclAllGather(send, recv, sendcount, dtype, self.comm, stream), "ncclAllGather")

    def reduce_scatter(self, send: int, recv: int, recvcount: int, dtype: int, stream: int) -> None:
        self._ck(self.lib.ncclReduceScatter(send, recv, recvcount, dtype, NCCL_SUM, self.comm, stream), "ncclReduceScatter")

    def broadcast(self, send: int, recv: int, count: int, dtype: int, root: int, stream: int) -> None:
        self._ck(self.lib.ncclBroadcast(send, recv, count, dtype, root, self.comm, stream), "ncclBroadcast")

    def reduce(self, send: int, recv: int, count: int, dtype: int, root: int, stream: int) -> None:
        self._ck(self.lib.ncclReduce(send, recv, count, dtype, NCCL_SUM, root, self.comm, stream), "ncclReduce")

    def send_recv(self, send: int, to: int, recv: int, frm: int, count: int, dtype: int, stream: int) -> None:
        # nccl-tests' sendrecv: one grouped send and recv (a ring step)
        self._ck(self.lib.ncclGroupStart(), "ncclGroupStart")
        self._ck(self.lib.ncclSend(send, count, dtype, to, self.comm, stream), "ncclSend")
        self._ck(self.lib.ncclRecv(recv, count, dtype, frm, self.comm, stream), "ncclRecv")
        self._ck(self.lib.ncclGroupEnd(), "ncclGroupEnd")

    def gather_scatter(self, op: str, send: int, recv: int, count: int, dtype: int, elem_size: int, root: int, stream: int) -> None:
        # nccl-tests' gather / scatter: grouped ncclSend / ncclRecv around the root (works on every NCCL that has send/recv)
        self._ck(self.lib.ncclGroupStart(), "ncclGroupStart")
        if op == "gather":
            self._ck(self.lib.ncclSend(send, count, dtype, root, self.comm, stream), "ncclSend")
            for p in (range(self.nranks) if self.rank == root else ()):
                self._ck(self.lib.ncclRecv(recv + p * count * elem_size, count, dtype, p, self.comm, stream), "ncclRecv")
        else:
            for p in (range(self.nranks) if self.rank == root else ()):
                self._ck(self.lib.ncclSend(send + p * count * elem_size, count, dtype, p, self.comm, stream), "ncclSend")
            self._ck(self.lib.ncclRecv(recv, count, dtype, root, self.comm, stream), "ncclRecv")
        self._ck(self.lib.ncclGroupEnd(), "ncclGroupEnd")

    def all_to_all(self, send: int, recv: int, count: int, dtype: int, elem_size: int, stream: int) -> None:
        # nccl-tests' alltoall: grouped ncclSend/ncclRecv to every peer
        self._ck(self.lib.ncclGroupStart(), "ncclGroupStart")
        for p in range(self.nranks):
            self._ck(self.lib.ncclSend(send + p * count * elem_size, count, dtype, p, self.comm, stream), "ncclSend")
            self._ck(self.lib.ncclRecv(recv + p * count * elem_size, count, dtype, p, self.comm, stream), "ncclRecv")
        self._ck(self.lib.ncclGroupEnd(), "ncclGroupEnd")

    def sym_tensor(self, numel: int, dtype, itemsize: int, windows: list):
        """A torch tensor on ncclMemAlloc memory registered as a symmetric window (NCCL_WIN_COLL_SYMMETRIC): NCCL 2.27+'s own
        zero-copy / symmetric-kernel path, i.e. the NCCL counterpart of our arena tensors. A collective call (same order on every rank)."""
        import torch
        nbytes = (numel * itemsize + 4095) // 4096 * 4096
        p = C.c_void_p()
        self._ck(self.lib.ncclMemAlloc(C.byref(p), nbytes), "ncclMemAlloc")
        win = C.c_void_p()
        self._ck(self.lib.ncclCommWindowRegister(self.comm, p, nbytes, C.byref(win), 0x01), "ncclCommWindowRegister")
        windows.append((win, p))

        class _Arr:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Arr(p.value, nbytes), device="cuda").view(dtype)[:numel]

    def destroy(self) -> None:
        if self.comm:
            self.lib.ncclCommDestroy(self.comm)
            self.comm = None
