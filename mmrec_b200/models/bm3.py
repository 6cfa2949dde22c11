"""BM3 on the B200 hot path; mirrors `/root/reference/src/models/bm3.py` (class name, config keys, parameter
names and order).  forward `:84-95` -> ops.propagate_mean; projections `:102-104` -> ops.project over the whole
table (the dropout targets `:110-119` draw a mask over the whole projected table, so the gather cannot move in
front of it without changing the RNG stream); full_sort_predict `:149-154` -> predictor + ops.score."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.functional import cosine_similarity

from .. import graph, ops
from ..common.abstract_recommender import GeneralRecommender
from ..common.loss import EmbLoss


class BM3(GeneralRecommender):
    def __init__(self, config, dataset):
        super().__init__(config, dataset)
        self.embedding_dim = config["embedding_size"]
        self.feat_embed_dim = config["embedding_size"]
        self.n_layers = config["n_layers"]
        self.reg_weight = config["reg_weight"]
        self.cl_weight = config["cl_weight"]
        self.dropout = config["dropout"]
        self.n_nodes = self.n_users + self.n_items
        self.norm_adj = graph.build_norm_adj(dataset.inter_matrix(form="coo").astype(np.float32), self.n_users,
                                             self.n_items, self.device)
        self.user_embedding = nn.Embedding(self.n_users, self.embedding_dim)
        self.item_id_embedding = nn.Embedding(self.n_items, self.embedding_dim)
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.item_id_embedding.weight)
        self.predictor = nn.Linear(self.embedding_dim, self.embedding_dim)
        self.reg_loss = EmbLoss()
        nn.init.xavier_normal_(self.predictor.weight)
        if self.v_feat is not None:
            self.image_embedding = nn.Embedding.from_pretrained(self.v_feat, freeze=False)
            self.image_trs = nn.Linear(self.v_feat.shape[1], self.feat_embed_dim)
            nn.init.xavier_normal_(self.image_trs.weight)
        if self.t_feat is not None:
            self.text_embedding = nn.Embedding.from_pretrained(self.t_feat, freeze=False)
            self.text_trs = nn.Linear(self.t_feat.shape[1], self.feat_embed_dim)
            nn.init.xavier_normal_(self.text_trs.weight)

    def forward(self):
        h = self.item_id_embedding.weight
        ego = torch.cat((self.user_embedding.weight, h), dim=0)
        all_emb = ops.propagate_mean(self.norm_adj, ego, self.n_layers)
        u_g, i_g = torch.split(all_emb, [self.n_users, self.n_items], dim=0)
        return u_g, i_g + h

    def calculate_loss(self, interactions):
        u_online_ori, i_online_ori = self.forward()
        t_feat_online, v_feat_online = None, None
        if self.t_feat is not None:
            t_feat_online = ops.project(self.text_embedding.weight, self.text_trs.weight, self.text_trs.bias)
        if self.v_feat is not None:
            v_feat_online = ops.project(self.image_embedding.weight, self.image_trs.weight, self.image_trs.bias)
        with torch.no_grad():
            u_target, i_target = u_online_ori.clone(), i_online_ori.clone()
            u_target = F.dropout(u_target, self.dropout)
            i_target = F.dropout(i_target, self.dropout)
            if self.t_feat is not None:
                t_feat_target = F.dropout(t_feat_online.clone(), self.dropout)
            if self.v_feat is not None:
                v_feat_target = F.dropout(v_feat_online.clone(), self.dropout)
        u_online, i_online = self.predictor(u_online_ori), self.predictor(i_online_ori)
        users, items = interactions[0], interactions[1]
        u_online, i_online = u_online[users, :], i_online[items, :]
        u_target, i_target = u_target[users, :], i_target[items, :]
        loss_t, loss_v, loss_tv, loss_vt = 0.0, 0.0, 0.0, 0.0
        if self.t_feat is not None:
            t_feat_online = self.predictor(t_feat_online)[items, :]
            t_feat_target = t_feat_target[items, :]
            loss_t = 1 - cosine_similarity(t_feat_online, i_target.detach(), dim=-1).mean()
            loss_tv = 1 - cosine_similarity(t_feat_online, t_feat_target.detach(), dim=-1).mean()
        if self.v_feat is not None:
            v_feat_online = self.predictor(v_feat_online)[items, :]
            v_feat_target = v_feat_target[items, :]
            loss_v = 1 - cosine_similarity(v_feat_online, i_target.detach(), dim=-1).mean()
            loss_vt = 1 - cosine_similarity(v_feat_online, v_feat_target.detach(), dim=-1).mean()
        loss_ui = 1 - cosine_similarity(u_online, i_target.detach(), dim=-1).mean()
        loss_iu = 1 - cosine_similarity(i_online, u_target.detach(), dim=-1).mean()
        return (loss_ui + loss_iu).mean() + self.reg_weight * self.reg_loss(u_online_ori, i_online_ori) + \
            self.cl_weight * (loss_t + loss_v + loss_tv + loss_vt).mean()

    def _score_embeddings(self):
        def run():
            u, i = self.forward()
            return self.predictor(u), self.predictor(i)
        return self._cached_eval_embeddings(run)

    def full_sort_predict(self, interaction):
        u, i = self._score_embeddings()
        return ops.score(u, i, interaction[0])
