"""What every two-domain estimator of the reference repeats verbatim: loader
construction (pygda/models/a2gnn.py:254-288 == udagcn.py:205-256 == grade.py:201-252 ==
adagcn.py:200-251), the epoch loop skeleton and ``predict`` (a2gnn.py:356-411)."""
import time

import torch

from ..data import DataLoader, NeighborLoader
from ..metrics import micro_f1_from_logits
from ..utils import logger


class TwoDomainLoop:
    # node-mode fit() on host-resident graphs: the loaders issue the NEXT epoch's host->device copy on a side stream
    # while the current step runs (data.NeighborLoader); same bytes, same values (tests/test_zz5_gpu_prefetch.py)
    prefetch = True

    def _build_loaders(self, source_data, target_data, prefetch=None):
        if self.mode == 'node':
            self.num_source_nodes, _ = source_data.x.shape
            self.num_target_nodes, _ = target_data.x.shape
            pin = str(self.device).startswith('cuda')       # host-resident graphs: pinned (packed) staging, once
            want = getattr(self, 'prefetch', False) if prefetch is None else prefetch
            pf = self.device if (pin and want) else None
            if self.batch_size == 0:
                self.source_batch_size = source_data.x.shape[0]
                self.source_loader = NeighborLoader(source_data, self.num_neigh,
                                                    batch_size=self.source_batch_size, pin=pin,
                                                    prefetch_device=pf)
                self.target_batch_size = target_data.x.shape[0]
                self.target_loader = NeighborLoader(target_data, self.num_neigh,
                                                    batch_size=self.target_batch_size, pin=pin,
                                                    prefetch_device=pf)
            else:
                self.source_loader = NeighborLoader(source_data, self.num_neigh, batch_size=self.batch_size, pin=pin,
                                                    prefetch_device=pf)
                self.target_loader = NeighborLoader(target_data, self.num_neigh, batch_size=self.batch_size, pin=pin,
                                                    prefetch_device=pf)
        elif self.mode == 'graph':
            # datasets given as sequences of graphs are made resident on the device once and collated there
            dev = self.device if isinstance(source_data, (list, tuple)) and isinstance(target_data, (list, tuple)) \
                else None
            if self.batch_size == 0:
                self.source_loader = DataLoader(source_data, batch_size=len(source_data), shuffle=True, device=dev)
                self.target_loader = DataLoader(target_data, batch_size=len(target_data), shuffle=True, device=dev)
            else:
                self.source_loader = DataLoader(source_data, batch_size=self.batch_size, shuffle=True, device=dev)
                self.target_loader = DataLoader(target_data, batch_size=self.batch_size, shuffle=True, device=dev)
        else:
            assert self.mode in ('graph', 'node'), 'Invalid train mode'

    def _fit_loop(self, step_fn):
        """``step_fn(epoch, source_batch, target_batch) -> (loss, source_logits, source_batch_on_device)``"""
        start_time = time.time()
        for epoch in range(self.epoch):
            epoch_loss = 0
            epoch_source_logits = None
            epoch_source_labels = None
            for idx, (sampled_source_data, sampled_target_data) in enumerate(
                    zip(self.source_loader, self.target_loader)):
                loss, source_logits, sampled_source_data = step_fn(epoch, sampled_source_data,
                                                                  sampled_target_data)
                epoch_loss += loss.item()
                if idx == 0:
                    epoch_source_logits, epoch_source_labels = source_logits, sampled_source_data.y
                else:
                    epoch_source_logits = torch.cat((epoch_source_logits, source_logits))
                    epoch_source_labels = torch.cat((epoch_source_labels, sampled_source_data.y))
            micro_f1_score = None
            if self.verbose > 1:
                # the reference scores F1 every epoch even when nothing is printed; the value is
                # only ever printed, so it is skipped for verbose <= 1 (no observable change)
                # eval_micro_f1(labels, logits.argmax(dim=1)) (a2gnn.py:328-329) with the argmax + counting on the GPU
                micro_f1_score = micro_f1_from_logits(epoch_source_labels, epoch_source_logits)
            logger(epoch=epoch, loss=epoch_loss, source_train_acc=micro_f1_score,
                   time=time.time() - start_time, verbose=self.verbose, train=True)

    def _predict_loop(self, logits_fn, source):
        """``predict`` ignores its ``data`` argument in the reference and re-iterates the loaders
        stored by ``fit``; kept."""
        loader = self.source_loader if source else self.target_loader
        logits = labels = None
        for idx, sampled_data in enumerate(loader):
            sampled_data = sampled_data.to(self.device)
            with torch.no_grad():
                out = logits_fn(sampled_data)
                if idx == 0:
                    logits, labels = out, sampled_data.y
                else:
                    logits = torch.cat((logits, out))
                    labels = torch.cat((labels, sampled_data.y))
        return logits, labels
