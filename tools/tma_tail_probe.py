"""Round-2 finding: a TMA tile::gather4 whose tensor map claims far more rows than exist (round 1 encoded 2^31 rows
"because only -1 must be out of range") faults with "warp illegal address" for some placements of the operand near the
end of a mapping (tiny operands in the last block of a caching-allocator segment; any operand that ends close to an
unmapped page).  With the true row count in the map it never does.  This probe places a [M, 64] bf16 operand `gap` bytes
before the end of a 2 MiB mapping that is followed by reserved-but-unmapped address space (driver VMM API) and runs
pnx_igemm on it: every line must say ok."""
import subprocess, sys
if len(sys.argv) > 1:
    import torch
    sys.path.insert(0, ".")
    from pillarnext_b200 import ops
    gap, M, which = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    # a 2 MiB mapping followed by 2 MiB of RESERVED BUT UNMAPPED address space (driver VMM API): any access past the end faults
    from cuda import cuda as cu
    torch.zeros(1, device="cuda")

    def ck(r):
        assert r[0] == cu.CUresult.CUDA_SUCCESS, r[0]
        return r[1] if len(r) > 1 else None

    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = 0
    gran = ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
    size = max(2 << 20, gran)
    va0 = ck(cu.cuMemAddressReserve(3 * size, 0, 0, 0))            # [unmapped | mapped | unmapped]
    va = cu.CUdeviceptr(int(va0) + size)
    hnd = ck(cu.cuMemCreate(size, prop, 0))
    ck(cu.cuMemMap(va, size, 0, hnd, 0))
    acc = cu.CUmemAccessDesc()
    acc.location = prop.location
    acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
    ck(cu.cuMemSetAccess(va, size, [acc], 1))

    class _Mem:
        __cuda_array_interface__ = {"shape": (size,), "typestr": "|u1", "data": (int(va), False), "version": 2}

    seg = torch.as_tensor(_Mem(), device="cuda")
    assert size == 2 << 20

    def at_end(nbytes):
        if which.endswith("_start"):
            return seg[gap:gap + nbytes]
        return seg[(2 << 20) - gap - nbytes:(2 << 20) - gap]

    x = torch.randn(M, 64, device="cuda").bfloat16()
    w = (torch.randn(9, 64, 64, device="cuda") * 0.1).bfloat16()
    nbr = torch.full((M, 9), -1, device="cuda", dtype=torch.int32)
    nbr[:, 4] = torch.arange(M, device="cuda", dtype=torch.int32)
    out = torch.empty(M, 64, device="cuda", dtype=torch.bfloat16)
    if which in ("x", "x_start"):
        x = at_end(M * 128).view(torch.bfloat16).view(M, 64).copy_(x)
    elif which == "w":
        w = at_end(9 * 64 * 128).view(torch.bfloat16).view(9, 64, 64).copy_(w)
    elif which == "nbr":
        nbr = at_end(M * 36).view(torch.int32).view(M, 9).copy_(nbr)
    elif which == "out":
        out = at_end(M * 128).view(torch.bfloat16).view(M, 64)
    ops.igemm(x, M, w, 9, 64, 64, out, nbr=nbr)
    torch.cuda.synchronize()
    print("ok")
else:
    for which in ("x",):
        for M in (3, 5, 8, 127, 129, 1001):
            for gap in (0, 1024, 32768):
                r = subprocess.run([sys.executable, __file__, str(gap), str(M), which], capture_output=True, text=True)
                print("%-3s M %4d  gap %5d B: %s" % (which, M, gap, "ok" if "ok" in r.stdout else "FAULT " + r.stderr[-120:].replace("\n", " ")))
