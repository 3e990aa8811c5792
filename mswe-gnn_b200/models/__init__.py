"""Model registry mirroring the reference's ``utils/miscellaneous.py:15-18`` (``get_model``)."""


def get_model(model_name):
    from .gnn import GNN, MSGNN
    models = {'GNN': GNN, 'MSGNN': MSGNN}
    return models[model_name]
