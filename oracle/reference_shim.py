"""TEST INFRASTRUCTURE ONLY -- loader for the *real* reference modules (authoring container only).

`/root/reference` is pure Python but does not import as shipped:
`gin` is not installed and `modules/h_rqvae.py:8` imports a misspelt loss class.  This module installs the
minimal shims described in SURVEY.md section 8(c) and returns the reference modules.  Nothing from the
reference is copied; it is imported in place.  The GPU box has no `/root/reference`, so only
`oracle/make_golden.py` and the `reference_available` tests call this.
"""
import enum
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HIDVAE_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "quantize.py"))


def _install_gin_stub() -> None:
    if "gin" in sys.modules:
        return
    gin = types.ModuleType("gin")

    def _identity_decorator(obj=None, *a, **k):
        if obj is None or not callable(obj) and not isinstance(obj, type):
            return lambda o: o
        return obj

    gin.configurable = _identity_decorator
    gin.constants_from_enum = _identity_decorator
    gin.parse_config_file = lambda *a, **k: None
    sys.modules["gin"] = gin


def load_reference():
    """Import the reference's RQ-path modules in place and return them as a namespace."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")  # h_rqvae.py:585 decorates forward with torch.compile
    _install_gin_stub()
    # The reference's top-level packages are called `modules`, `init`, `data`, `distributions`; make sure
    # no same-named package of ours shadows them in this process.
    for name in list(sys.modules):
        root = name.split(".")[0]
        if root in ("modules", "init", "data", "distributions"):
            mod = sys.modules[name]
            f = getattr(mod, "__file__", "") or ""
            if not f.startswith(REFERENCE_ROOT):
                del sys.modules[name]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import modules.loss as ref_loss  # noqa: E402
    if not hasattr(ref_loss, "CategoricalReconstuctionLoss"):
        ref_loss.CategoricalReconstuctionLoss = ref_loss.CategoricalReconstructionLoss  # h_rqvae.py:8 typo
    import modules.quantize as ref_quantize  # noqa: E402
    import modules.h_rqvae as ref_h_rqvae  # noqa: E402
    import init.kmeans as ref_kmeans  # noqa: E402
    import data.schemas as ref_schemas  # noqa: E402
    ns = types.SimpleNamespace(
        loss=ref_loss, quantize=ref_quantize, h_rqvae=ref_h_rqvae, kmeans=ref_kmeans, schemas=ref_schemas
    )
    return ns


def load_reference_tokenizer():
    """The reference's `modules/tokenizer/h_semids.py`.  It imports `data.tags_processed` for three dataset class
    names only (that module needs polars / torch_geometric, which are not installed): a stub module with those names
    stands in, everything the tokenizer itself computes is the reference's own code."""
    load_reference()
    if "data.tags_processed" not in sys.modules:
        stub = types.ModuleType("data.tags_processed")
        for name in ("ItemData", "SeqData", "RecDataset"):
            setattr(stub, name, type(name, (), {}))
        sys.modules["data.tags_processed"] = stub
    import modules.tokenizer.h_semids as ref_tok  # noqa: E402
    return ref_tok
