"""oracle/ — CPU restatement of the reference's captioning hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the shipped package (``imagecaptioningconvnext_b200``) imports this.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and
only as the checker or as the reported CPU baseline — never as the thing shipped.

The reference (sa06840/ImageCaptioningConvNeXt) is pure Python on top of torch / torchvision; its arithmetic
lives in those two third-party packages (torch 2.11.0, torchvision 0.26.0 in this image; the reference pins
neither).  Every function here restates one reference function in plain ``torch.nn.functional`` ops on CPU
tensors (fp32 by default, fp64 on request for tie labelling) and cites the file:line it follows.

Pinning: the reference ships no tests or golden vectors.  The oracle is therefore pinned
against outputs of the *reference itself*, imported in the build container from /root/reference with three
shims (gensim stub, no-download convnext_base, matplotlib/skimage stubs) by ``tests/golden/make_golden.py``;
the resulting small fixtures live in ``tests/golden/*.pt`` and ``tests/test_oracle_golden.py`` checks the oracle
against them on every CPU run.
"""
