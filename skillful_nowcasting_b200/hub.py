"""PyTorchModelHubMixin with a save path that works while the parameters are views of an optimiser's flat buffer.

`training.Adam` re-points every parameter to a slice of ONE flat tensor (fused Adam launch, one all-reduce).  safetensors refuses to
serialise tensors that share a larger storage ("None is covering the entire storage"), which is what the stock
`PyTorchModelHubMixin._save_pretrained` -> `save_model_as_safetensor` then hits.  Saving compact copies keeps the on-disk format (and
`from_pretrained`) exactly the reference's: `model.safetensors` with the reference's state-dict keys + `config.json`
(ref: README.md:57-69, tests/test_model.py:341-399)."""
from __future__ import annotations

from pathlib import Path

from huggingface_hub import PyTorchModelHubMixin


class HubMixin(PyTorchModelHubMixin):
    def _save_pretrained(self, save_directory: Path) -> None:
        from safetensors.torch import save_file

        model = self.module if hasattr(self, "module") else self
        state = {k: v.detach().contiguous().clone().cpu() for k, v in model.state_dict().items()}
        save_file(state, str(Path(save_directory) / "model.safetensors"), metadata={"format": "pt"})
