"""Bilevel runner with the reference's API (src/trainers/bilevel.py:16-145). Host control flow only: the
early-stopped inner loop, a hypergradient step every `hyper_gradient_interval` inner steps (the hot path,
`hyper_opt_step` -> `OuterProblemTrainer.train_step`), the empirical-mean evaluation and the in-memory
checkpoint of the best (GCN weights, graph-model state_dict)."""
import os
from copy import deepcopy
from typing import Dict

from ..utils.early_stopping import EarlyStopping
from ..utils.evaluation import empirical_mean_loss
from ..utils.graph import DenseData
from ..utils.tracking import setup_basic_logger
from . import Metrics
from .inner import InnerProblemTrainer
from .outer import OuterProblemTrainer


class BilevelProblemRunner:

    def __init__(self, inner_trainer: InnerProblemTrainer, outer_trainer: OuterProblemTrainer, data: DenseData,
                 n_samples_empirical_mean: int = 16):
        self.inner_trainer = inner_trainer
        self.outer_trainer = outer_trainer
        self.data = data
        self.gcn_params = None
        self.graph_state_dict = None
        self.n_samples_empirical_mean = n_samples_empirical_mean
        self.logger = setup_basic_logger()
        # Replay aligned blocks (tau inner steps + the hyper step) from one captured CUDA graph when the plain LDS configuration
        # allows it (trainers/graph_block.py); LDS_GRAPH_BLOCKS=0 or `graph_blocks = False` keeps the step-by-step loop.
        self.graph_blocks = os.environ.get("LDS_GRAPH_BLOCKS", "1") != "0"
        # The reference keeps `graph_model.state_dict()` WITHOUT a copy in its early stopper (src/trainers/bilevel.py:96-98): the
        # dict aliases the live `probs`, so `evaluate()` reloads the LAST theta next to the best GCN weights. False (default)
        # reproduces that; True snapshots theta of the best outer iteration instead (a deliberate deviation).
        self.snapshot_best_graph = False
        self._blocks = {}

    def train(self, patience: int, hyper_gradient_interval: int, inner_loop_max_epochs: int = 400,
              outer_loop_max_epochs: int = 400, sacred_runner=None):
        outer_stopper = EarlyStopping(patience=patience, max_epochs=outer_loop_max_epochs)
        current_step = 0
        outer_step = 0
        while not outer_stopper.abort:                         # judged on the empirical-mean validation loss
            inner_stopper = EarlyStopping(patience=patience, max_epochs=inner_loop_max_epochs)
            self.inner_trainer.reset_weights()
            self.inner_trainer.reset_optimizer()
            self.logger.info("Starting new outer loop...")
            tau = max(1, hyper_gradient_interval)
            while not inner_stopper.abort:                     # judged on the training loss
                if self.graph_blocks and tau <= 64 and current_step % tau == 1 % tau and self._block_eligible():
                    advanced = self._run_block(tau, current_step, inner_stopper, sacred_runner)
                    if advanced is not None:
                        current_step = advanced
                        continue
                train_metrics = self.inner_opt_step()
                inner_stopper.update(train_metrics.loss, model_params=self.inner_trainer.copy_model_params())
                if sacred_runner is not None:
                    sacred_runner.log_scalar("loss.train", train_metrics.loss, step=current_step)
                    sacred_runner.log_scalar("acc.train", train_metrics.acc, step=current_step)
                    sacred_runner.log_scalar("Memory Usage (%)", _memory_percent())
                self.logger.info(f"Model Optimization Step {current_step}: loss={train_metrics.loss}, accuracy={train_metrics.acc}")
                if hyper_gradient_interval == 0 or current_step % hyper_gradient_interval == 0:
                    self.hyper_opt_step(current_step, sacred_runner)
                current_step += 1
            self.logger.info("Exited inner optimization")
            gcn_params = inner_stopper.model_params
            self.outer_trainer.train(False)
            val_results, test_results = empirical_mean_loss(self.inner_trainer.model, graph_model=self.outer_trainer.model,
                                                            n_samples=self.n_samples_empirical_mean, data=self.data,
                                                            model_parameters=gcn_params)
            if sacred_runner is not None:
                sacred_runner.log_scalar("loss.val.empirical", val_results.loss)
                sacred_runner.log_scalar("acc.val.empirical", val_results.acc)
                sacred_runner.log_scalar("loss.test.empirical", test_results.loss)
                sacred_runner.log_scalar("acc.test.empirical", test_results.acc)
            self.logger.info(f"Empirical Validation Set Results: loss={val_results.loss}, accuracy={val_results.acc}")
            graph_state = self.outer_trainer.model.state_dict()
            outer_stopper.update(val_results.loss, model_params=[deepcopy(gcn_params), deepcopy(graph_state) if self.snapshot_best_graph else graph_state])
            outer_step += 1
        self.logger.info(f"Ended training after {outer_step} steps...")
        self.gcn_params, self.graph_state_dict = outer_stopper.model_params

    # ---- captured blocks ---------------------------------------------------------------------------
    def _block_eligible(self) -> bool:
        from .graph_block import CapturedBilevelBlock
        return CapturedBilevelBlock.eligible(self)

    def _run_block(self, tau: int, current_step: int, inner_stopper: EarlyStopping, sacred_runner) -> int:
        """Steps current_step .. current_step + tau - 1 of the reference loop (the last one triggers the hyper step) from ONE
        graph replay; the per-step early stopping is applied afterwards, and if the reference would have stopped before the
        block's hyper step, that step is undone. Returns the new current_step (None: capture failed, nothing was run)."""
        from .graph_block import CapturedBilevelBlock
        from .graph_block import BlockCaptureError
        block = self._blocks.get(tau)
        if block is None:
            block = self._blocks[tau] = CapturedBilevelBlock(self, tau)
        try:
            metrics = block.replay()
        except BlockCaptureError as exc:                       # state is back at block entry: carry on step by step
            self.logger.warning(f"bilevel block could not be captured into a CUDA graph ({exc}); continuing step by step")
            self.graph_blocks = False
            return None
        steps_done, last_kept = 0, None
        for k in range(tau):
            train_metrics = metrics[k]
            token = ("weights after step", k)                  # resolved to real tensors once, after the loop
            inner_stopper.update(train_metrics.loss, model_params=token)
            if inner_stopper.model_params is token:            # the stopper kept this step's weights (early_stopping.py:26-30)
                last_kept = k
            if sacred_runner is not None:
                sacred_runner.log_scalar("loss.train", train_metrics.loss, step=current_step)
                sacred_runner.log_scalar("acc.train", train_metrics.acc, step=current_step)
                sacred_runner.log_scalar("Memory Usage (%)", _memory_percent())
            self.logger.info(f"Model Optimization Step {current_step}: loss={train_metrics.loss}, accuracy={train_metrics.acc}")
            current_step += 1
            steps_done = k + 1
            if inner_stopper.abort:
                break
        if last_kept is not None:
            inner_stopper.model_params = block.params_after(last_kept)
        if steps_done == tau:                                  # the block's last step is the one that triggers the hyper step
            self._log_hyper_step(metrics[tau], current_step - 1, sacred_runner)
        else:
            block.undo_hyper_step(steps_done)
        if inner_stopper.abort:
            block.store_state(steps_done)
        return current_step

    def inner_opt_step(self) -> Metrics:
        self.outer_trainer.train()
        sampler = getattr(self.outer_trainer, "sample_for_unroll", None)       # factored graph when the hyper step can use it
        graph = sampler(self.inner_trainer) if sampler is not None else self.outer_trainer.sample()
        return self.inner_trainer.train_step(graph)

    def hyper_opt_step(self, current_step: int, sacred_runner=None):
        self.logger.info(f"Optimizing graph parameters at step {current_step}")
        metrics = self.outer_trainer.train_step(self.inner_trainer.model_forward)
        self.inner_trainer.detach()
        self.outer_trainer.detach()
        self._log_hyper_step(metrics, current_step, sacred_runner)

    def _log_hyper_step(self, metrics: Metrics, current_step: int, sacred_runner=None):
        if sacred_runner is not None:
            sacred_runner.log_scalar("loss.outer", metrics.loss, step=current_step)
            sacred_runner.log_scalar("acc.outer", metrics.acc, step=current_step)
            for i, lr in enumerate(self.outer_trainer.get_learning_rates()):
                sacred_runner.log_scalar(f"Outer Learning Rate {i}", lr, step=current_step)
            for name, value in self.outer_trainer.model.statistics().items():
                sacred_runner.log_scalar(name, value, step=current_step)
                self.logger.info(f"{name}: {value}")
        self.logger.info(f"Performance on held-out sample for graph optimization: loss={metrics.loss}, accuracy={metrics.acc}"
                         f"Outer optimizer learning rate: {self.outer_trainer.get_learning_rates()}")

    def evaluate(self) -> Dict:
        assert self.gcn_params is not None and self.graph_state_dict is not None, "Models need to be trained before evaluation."
        self.outer_trainer.model.load_state_dict(self.graph_state_dict)
        val_results, test_results = empirical_mean_loss(self.inner_trainer.model, graph_model=self.outer_trainer.model,
                                                        n_samples=self.n_samples_empirical_mean, data=self.data,
                                                        model_parameters=self.gcn_params)
        return {"loss.val.final": val_results.loss, "acc.val.final": val_results.acc,
                "loss.test.final": test_results.loss, "acc.test.final": test_results.acc}


def _memory_percent():
    try:
        import psutil
        return psutil.Process(os.getpid()).memory_percent()
    except Exception:
        return float("nan")
