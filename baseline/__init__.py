"""The reference arm: harness around the reference's OWN code (baseline/_ref, built by tools/make_ref.py).  Test and
benchmark infrastructure only -- nothing under incomplete_multimodal_fusion_b200/ imports it."""
