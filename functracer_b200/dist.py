"""One process per GPU: the gather step of the multi-GPU path.  Rendering needs no
collective; finished shard buffers are gathered to rank 0 (NCCL over NVLink on the GPU box, gloo in the
CPU tests) where the frame is assembled."""
import torch
import torch.distributed as dist


def shard_of_rank():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def gather_tiles(local_buf, max_elems, gather_list=None, dst=0):
    """Gathers every rank's tile buffer (padded to max_elems so that all contributions have one size) to
    `dst`.  local_buf: 1-D tensor of >= max_elems elements on the rank's device.  Returns the list of
    per-rank tensors on dst, None elsewhere."""
    rank, world = shard_of_rank()
    if world == 1:
        return [local_buf]
    send = local_buf[:max_elems]
    if rank == dst:
        if gather_list is None:
            gather_list = [torch.empty_like(send) for _ in range(world)]
        dist.gather(send, gather_list, dst=dst)
        return gather_list
    dist.gather(send, None, dst=dst)
    return None
