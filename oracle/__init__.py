"""CPU oracle for the EvoGP hot path — TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  ``evogp_b200`` never does.

``oracle.lib()``      -> ctypes handle of the C restatement (evogp_oracle.c)
``oracle.ref_gpu()``  -> ctypes handle of the reference's own CUDA kernels,
                         compiled unmodified from /root/reference by
                         ``oracle/build_ref.sh`` into ``oracle/_ref/`` (GPU only)
"""
from .oracle import (  # noqa: F401
    build,
    build_ref,
    build_reference_package,
    lib,
    ref_gpu,
    ref_gpu_available,
    evaluate,
    batch_forward,
    sr_fitness,
    generate,
    generate_philox,
    philox,
    extract_subtree,
    tournament,
    feistel_perm,
    next_generation,
    crossover,
    mutate,
    hash32,
    taus88_draws,
    taus88_nth,
    max_threads,
    check_forest,
)
