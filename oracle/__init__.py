"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (plain PyTorch-CPU / numpy) of the residual-quantisation hot path of HiD-VAE, used only as
the checker for the CUDA path: by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s `cpu_baseline`
/ `--impl reference` legs.  Nothing under `hid-vae_b200/` imports it and the product path never routes
through it (the product fails loudly when the CUDA library is missing).

Parity pinning: the reference repository ships no tests, golden vectors or fixtures, so
the oracle is pinned against *outputs of the reference itself*: `oracle/make_golden.py` imports the real
reference modules in the authoring container (through `oracle/reference_shim.py`) and writes the fixtures in
`tests/golden/`; `tests/test_oracle_golden.py` checks the oracle against them everywhere, and
`tests/test_oracle_vs_reference.py` checks it live against the reference wherever `/root/reference` exists.
"""
