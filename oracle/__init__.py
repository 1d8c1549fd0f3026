"""CPU oracle for the AnySD denoising hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and there only as the checker (or as the timed
CPU baseline), never as the thing shipped.  The product path
(``anyedit_b200``) never imports this package and fails loudly when its CUDA
library is missing.

What is here
------------
``unet_oracle``   fp32 torch-CPU restatement of ``ldm`` ``UNetModel.forward``
                  (openaimodel.py:754-786, attention.py:145-340, util.py:154-219)
``ddim_oracle``   numpy/torch restatement of the schedule helpers and the
                  ``DDIMSampler`` loop (util.py:21-74, ddim.py:23-251,
                  ddpm.py:138-166, 356-359, 1332-1363)
``anysd_oracle``  builder-written restatement of the task router / task
                  embedding (source absent from the reference: PARITY UNPINNED)
``weights``       deterministic, torch-RNG-independent weight generator
``ref_import``    imports the real reference from /root/reference (this
                  container only) -- used by tests/golden/make_golden.py

Pinning: the reference holds no golden vectors for this path, so
the restatement is pinned against outputs of the reference's own modules run in
the build container; those outputs are committed under ``tests/golden`` together
with the generating script (``tests/golden/make_golden.py``).
"""
