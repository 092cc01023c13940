"""`deepspeed.comm` stand-in (the reference imports it as `dist` in utils/common.py:9, utils/dataset.py:17-18,
utils/saver.py:8-9): thin pass-through to torch.distributed with the same call names."""
import os

import torch
import torch.distributed as tdist

ReduceOp = tdist.ReduceOp
tdist = tdist


def init_distributed(dist_backend=None, timeout=None):
    """deepspeed.init_distributed(): env-driven (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT); single process if unset."""
    if tdist.is_initialized():
        return
    if dist_backend is None:
        dist_backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    kw = {}
    if timeout is not None:
        kw['timeout'] = timeout
    if dist_backend == 'nccl':
        # NCCL for device collectives plus gloo for host-side object / metadata traffic
        dist_backend = 'cpu:gloo,cuda:nccl'
    tdist.init_process_group(backend=dist_backend, rank=rank, world_size=world, **kw)


def is_initialized():
    return tdist.is_initialized()


def get_rank(group=None):
    return tdist.get_rank(group) if tdist.is_initialized() else 0


def get_world_size(group=None):
    return tdist.get_world_size(group) if tdist.is_initialized() else 1


def get_world_group():
    return tdist.group.WORLD


def barrier(group=None):
    if tdist.is_initialized():
        tdist.barrier(group=group)


def new_group(ranks):
    return tdist.new_group(ranks=ranks)


def send(tensor, dst, group=None, tag=0):
    return tdist.send(tensor, dst, group=group, tag=tag)


def recv(tensor, src=None, group=None, tag=0):
    return tdist.recv(tensor, src=src, group=group, tag=tag)


def broadcast(tensor, src, group=None):
    return tdist.broadcast(tensor, src, group=group)


def all_reduce(tensor, op=ReduceOp.SUM, group=None):
    return tdist.all_reduce(tensor, op=op, group=group)


def all_gather_object(object_list, obj, group=None):
    if not tdist.is_initialized():          # single process (train.py only initialises torch.distributed when WORLD_SIZE > 1)
        object_list[0] = obj
        return None
    return tdist.all_gather_object(object_list, obj, group=group)


def broadcast_object_list(object_list, src=0, group=None):
    if not tdist.is_initialized():
        return None
    return tdist.broadcast_object_list(object_list, src=src, group=group)
