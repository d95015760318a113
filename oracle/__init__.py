"""CPU oracle for the CSR SpMV hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
reported CPU baseline.  The product (``spmv_samples_b200``) never imports it.

* ``oracle.cpu``        ctypes bindings of ``liboracle.so`` (the C restatement,
                        ``spmv_oracle.c``) and, when present, of
                        ``_ref/libspmv_ref.so`` (the reference's own CPU code compiled
                        from /root/reference by ``oracle/Makefile``).
* ``oracle.generators`` numpy restatement of the device matrix generators
                        (``spmv_samples_b200/csrc/gen.cu``), bit-for-bit.
"""
