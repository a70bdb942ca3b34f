"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (fp32 torch / numpy / plain C) of the reference algorithm for the
PillarNeXt-B hot path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import, link or execute
anything in here, and only as the checker or the CPU baseline -- never as the product path.
The product (``pillarnext_b200``) never imports this package.

Parity pinning: the reference repository has NO tests and NO golden vectors (SURVEY.md
section 4), so parity is pinned the only way possible: the restatement is checked, in the
build container, against the reference's OWN files executed on CPU
(``oracle/reference_loader.py``: reader, neck, head, loss imported unmodified from
/root/reference with a 2-function torch_scatter shim) and the resulting input/output
vectors are frozen under ``tests/golden/`` by ``oracle/make_golden.py``.
The sparse backbone's arithmetic lives in spconv (third party, unpinned, source absent):
for that stage parity is UNPINNED by any reference artefact; two independent restatements
(dense-masked and gather-GEMM-scatter) are required to agree instead.
"""
