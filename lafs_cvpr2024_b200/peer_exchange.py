"""Peer-memory exchange buffers for the class-sharded margin head (csrc/exchange.cu).

One symmetric buffer per (process group, B, D): allocated with torch.distributed._symmetric_memory
(the same virtual layout on every rank of one NVLink / NVSwitch domain), zero-filled, and described to
the kernels by the device-side table of peer base addresses.  The two exchange steps of a head step --
per-row softmax statistics and the sum of the partial embedding gradient -- then run as one kernel each
(lafs_xchg_stats / lafs_xchg_allreduce) instead of NCCL all_gather / all_reduce.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class PeerExchange:
    def __init__(self, group, B, D, device):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.B, self.D = B, D
        offs = (C.c_size_t * 5)()
        nbytes = _lib.lib().lafs_xchg_bytes(self.world, B, D, offs)
        if nbytes == 0:
            raise ValueError(f"unsupported exchange shape world={self.world} B={B} D={D} (<= 8 ranks, D % 4 == 0)")
        self.buf = symm.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.peer_table = int(self.hdl.buffer_ptrs_dev)          # device array of `world` base addresses
        o_in, o_out, o_err = offs[2] // 4, offs[3] // 4, offs[4] // 4
        self.de_in = self.buf[o_in:o_in + B * D].view(B, D)       # this rank's partial dE_hat
        self.de_out = self.buf[o_out:o_out + B * D].view(B, D)    # sum over the ranks
        self._err = self.buf[o_err:o_err + 1].view(torch.int32)
        torch.cuda.synchronize(device)
        dist.barrier(self.group)                                  # every rank's flags are zero before the first put

    def merge_stats(self, local_stats):
        """[B,4] records of this rank -> records merged over all ranks (every rank gets the same bits)."""
        merged = torch.empty_like(local_stats)
        _lib.call("lafs_xchg_stats", self.peer_table, self.rank, self.world, self.B, self.D, local_stats.data_ptr(),
                  merged.data_ptr(), _lib.stream())
        return merged

    def allreduce_de(self):
        """de_in of every rank summed into de_out of every rank."""
        _lib.call("lafs_xchg_allreduce", self.peer_table, self.rank, self.world, self.B, self.D, _lib.stream())
        return self.de_out

    def check(self):
        """Raises if a peer failed to arrive within the kernels' spin bound (synchronises the device)."""
        if int(self._err.item()) != 0:
            raise RuntimeError("lafs peer exchange: a rank did not arrive (spin bound exceeded)")
