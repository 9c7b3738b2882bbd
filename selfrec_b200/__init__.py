"""selfrec_b200: the B200-native hot path behind SELFRec's plugin surface.

    import selfrec_b200
    selfrec_b200.install()          # alias base.*, util.*, data.*, model.graph.* in sys.modules

after which the reference launcher (`SELFRec(conf).execute()`) and the reference's own
model files import this package's drop-in modules.  See INTEGRATION.md.
"""
import importlib
import sys

__version__ = "0.1.0"

_DROPIN = {
    # the five boundary modules (BASELINE.json north_star / SURVEY 8b)
    "base.graph_recommender": "selfrec_b200.base.graph_recommender",
    "base.torch_interface": "selfrec_b200.base.torch_interface",
    "util.loss_torch": "selfrec_b200.util.loss_torch",
    "util.sampler": "selfrec_b200.util.sampler",
    "data.ui_graph": "selfrec_b200.data.ui_graph",
    # file -> CSR through the native builder, ranking metrics from device hit masks
    "data.loader": "selfrec_b200.data.loader",
    "util.evaluation": "selfrec_b200.util.evaluation",
}
_FUSED_MODELS = {f"model.graph.{m}": f"selfrec_b200.model.graph.{m}" for m in ("MF", "LightGCN", "SimGCL", "XSimGCL", "SGL")}


def install(fused_models=True):
    """Register the drop-in modules under the reference's import names.

    With fused_models=True the five in-scope model classes resolve to the fused-engine
    versions too; with False the reference's own model files run on top of the five
    boundary modules (op-level drop-in)."""
    table = dict(_DROPIN)
    if fused_models:
        table.update(_FUSED_MODELS)
    for alias, target in table.items():
        sys.modules[alias] = importlib.import_module(target)
    return sorted(table)
