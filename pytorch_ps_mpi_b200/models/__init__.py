"""Model zoo for the BASELINE.json configurations (random-init, synthetic data; no downloads)."""
from .mlp import MLP, mnist_mlp
from .resnet import ResNet, resnet18, resnet50
from .bert import BertConfig, BertModel, BertForPreTraining, bert_base

__all__ = ["MLP", "mnist_mlp", "ResNet", "resnet18", "resnet50", "BertConfig", "BertModel",
           "BertForPreTraining", "bert_base", "build"]


def build(name: str, **kw):
    """``build('resnet18' | 'resnet50' | 'bert_base' | 'mlp')``."""
    table = {"mlp": mnist_mlp, "resnet18": resnet18, "resnet50": resnet50, "bert_base": bert_base}
    return table[name](**kw)
