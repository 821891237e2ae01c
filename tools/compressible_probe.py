#!/usr/bin/env python
"""Probe: does an observation tensor in COMPRESSIBLE device memory (cuMemCreate with
CU_MEM_ALLOCATION_COMP_GENERIC — the L2 compresses lines on their way to HBM) shorten the rasteriser?
The observation maps are mostly zeros (a few dozen non-zero cells of 3,072 per env), the best case for
the hardware compressor.  Prints one JSON object; not part of the product path.

    python tools/compressible_probe.py [--envs 65536] [--reps 30]
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rad_team_b200.envs import BatchedRadSearchEnv  # noqa: E402


def ck(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"driver call failed: {err}")
    return res[1] if len(res) == 2 else res[1:]


def alloc_compressible(nbytes, device=0):
    from cuda.bindings import driver as cu

    ck(cu.cuInit(0))
    dev = ck(cu.cuDeviceGet(device))
    sup = ck(cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_GENERIC_COMPRESSION_SUPPORTED, dev))
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = device
    prop.allocFlags.compressionType = 1  # CU_MEM_ALLOCATION_COMP_GENERIC
    gran = ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED))
    size = (nbytes + gran - 1) // gran * gran
    handle = ck(cu.cuMemCreate(size, prop, 0))
    got = ck(cu.cuMemGetAllocationPropertiesFromHandle(handle))
    ptr = ck(cu.cuMemAddressReserve(size, 0, 0, 0))
    ck(cu.cuMemMap(ptr, size, 0, handle, 0))
    acc = cu.CUmemAccessDesc()
    acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    acc.location.id = device
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(cu.cuMemSetAccess(ptr, size, [acc], 1))
    return int(ptr), size, int(sup), int(got.allocFlags.compressionType)


def as_tensor(ptr, shape, typestr="<f4"):
    class _A:
        pass

    a = _A()
    a.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2,
                                  "strides": None}
    return torch.as_tensor(a, device="cuda:0")


def timeit(fn, reps, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=30)
    args = ap.parse_args()
    torch.zeros(1, device="cuda:0")
    E, A, G = args.envs, 3, 32
    n = E * 3 * G * G
    out = {"envs": E}
    ptr, size, sup, got = alloc_compressible(n * 4)
    out.update(compression_supported=sup, compression_type_granted=got, bytes=size)
    comp = as_tensor(ptr, (E, 3, G, G))
    plain = torch.empty((E, 3, G, G), dtype=torch.float32, device="cuda:0")

    env = BatchedRadSearchEnv(n_envs=E, n_agents=A, grid=G, max_steps=120, seed=1, alloc_maps=False)
    env.reset(out_maps=plain)
    acts = torch.randint(1, 5, (E, A), dtype=torch.int32, device="cuda:0")
    for _ in range(60):  # mid-episode: ~100 visited cells per env
        env.step(acts.copy_(torch.randint(1, 5, (E, A), dtype=torch.int32, device="cuda:0")), rasterize=False)
    env.rasterize(plain)
    env.rasterize(comp)
    torch.cuda.synchronize()
    out["equal"] = bool(torch.equal(plain, comp))
    gb = n * 4 / 1e9
    for name, t in (("plain", plain), ("compressible", comp)):
        ms = timeit(lambda: env.rasterize(t), args.reps)
        out[f"raster_{name}_ms"] = ms
        out[f"raster_{name}_gbs"] = gb / ms * 1e3
        ms = timeit(lambda: t.zero_(), args.reps)
        out[f"fill0_{name}_ms"] = ms
        out[f"fill0_{name}_gbs"] = gb / ms * 1e3
        ms = timeit(lambda: t.sum(), args.reps)
        out[f"sum_{name}_ms"] = ms
        out[f"sum_{name}_gbs"] = gb / ms * 1e3
        env.rasterize(t)
    # consumer-side view: a reduction over the rasterised (mostly zero) tensor
    for name, t in (("plain", plain), ("compressible", comp)):
        ms = timeit(lambda: t.sum(), args.reps)
        out[f"sum_rasterised_{name}_ms"] = ms
    print(json.dumps(out))


if __name__ == "__main__":
    main()
