"""
oracle/ — TEST INFRASTRUCTURE, not product code.

CPU restatement (numpy) of the reference algorithm for the hot path
``HarmonizeRecord -> TRRecord.GetLengthGenotypes/GetAlleleCounts/GetGenotypeCounts
-> {statSTR stats | dumpSTR filters | associaTR OLS}`` of gymrek-lab/TRTools
(reference @ f8ef1e9, v6.1.0).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package ``trtools_b200`` never does (it fails loudly without its CUDA
library).  See oracle/README.md for how the restatement is pinned.
"""
