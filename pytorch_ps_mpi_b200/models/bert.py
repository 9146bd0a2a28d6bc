"""BERT-base encoder with MLM + NSP heads (BASELINE.json config 4), bf16-friendly.

Standard post-LN BERT (Devlin et al. 2018): 12 layers, hidden 768, 12 heads, FFN 3072,
vocab 30522, 512 positions → ≈110 M parameters.  Attention goes through
``F.scaled_dot_product_attention`` (library flash kernels); the framework's own kernels are on
the gradient / parameter paths (top-k encode, PS gather-update, broadcast GEMM).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    max_position_embeddings: int = 512
    type_vocab_size: int = 2
    layer_norm_eps: float = 1e-12
    dropout: float = 0.0


class BertEmbeddings(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.word_embeddings = nn.Embedding(c.vocab_size, c.hidden_size)
        self.position_embeddings = nn.Embedding(c.max_position_embeddings, c.hidden_size)
        self.token_type_embeddings = nn.Embedding(c.type_vocab_size, c.hidden_size)
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.dropout = nn.Dropout(c.dropout)

    def forward(self, input_ids, token_type_ids=None):
        B, S = input_ids.shape
        pos = torch.arange(S, device=input_ids.device)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        x = self.word_embeddings(input_ids) + self.position_embeddings(pos)[None] + \
            self.token_type_embeddings(token_type_ids)
        return self.dropout(self.LayerNorm(x))


class BertLayer(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.nh = c.num_attention_heads
        self.qkv = nn.Linear(c.hidden_size, 3 * c.hidden_size)
        self.attn_out = nn.Linear(c.hidden_size, c.hidden_size)
        self.ln1 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.ffn_in = nn.Linear(c.hidden_size, c.intermediate_size)
        self.ffn_out = nn.Linear(c.intermediate_size, c.hidden_size)
        self.ln2 = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.drop = nn.Dropout(c.dropout)

    def forward(self, x, attn_mask=None):
        B, S, H = x.shape
        q, k, v = self.qkv(x).view(B, S, 3, self.nh, H // self.nh).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask)
        a = a.transpose(1, 2).reshape(B, S, H)
        x = self.ln1(x + self.drop(self.attn_out(a)))
        h = self.ffn_out(F.gelu(self.ffn_in(x)))
        return self.ln2(x + self.drop(h))


class BertModel(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.config = c
        self.embeddings = BertEmbeddings(c)
        self.layers = nn.ModuleList(BertLayer(c) for _ in range(c.num_hidden_layers))
        self.pooler = nn.Linear(c.hidden_size, c.hidden_size)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None):
        x = self.embeddings(input_ids, token_type_ids)
        m = None
        if attention_mask is not None:
            m = attention_mask[:, None, None, :].to(torch.bool)
        for layer in self.layers:
            x = layer(x, m)
        return x, torch.tanh(self.pooler(x[:, 0]))


class BertForPreTraining(nn.Module):
    """MLM (tied decoder) + NSP heads; ``forward`` returns the summed loss when labels are given."""

    def __init__(self, c: BertConfig):
        super().__init__()
        self.bert = BertModel(c)
        self.transform = nn.Linear(c.hidden_size, c.hidden_size)
        self.transform_ln = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)
        self.decoder_bias = nn.Parameter(torch.zeros(c.vocab_size))
        self.nsp = nn.Linear(c.hidden_size, 2)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        if isinstance(m, (nn.Linear, nn.Embedding)):
            nn.init.normal_(m.weight, std=0.02)
            if isinstance(m, nn.Linear) and m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, input_ids, token_type_ids=None, attention_mask=None, mlm_labels=None, nsp_labels=None):
        seq, pooled = self.bert(input_ids, token_type_ids, attention_mask)
        h = self.transform_ln(F.gelu(self.transform(seq)))
        logits = F.linear(h, self.bert.embeddings.word_embeddings.weight, self.decoder_bias)
        nsp_logits = self.nsp(pooled)
        if mlm_labels is None:
            return logits, nsp_logits
        loss = F.cross_entropy(logits.view(-1, logits.size(-1)).float(), mlm_labels.view(-1), ignore_index=-100)
        if nsp_labels is not None:
            loss = loss + F.cross_entropy(nsp_logits.float(), nsp_labels)
        return loss


def bert_base(**kw) -> BertForPreTraining:
    return BertForPreTraining(BertConfig(**kw))
