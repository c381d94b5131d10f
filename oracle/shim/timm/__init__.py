"""Minimal stand-in for timm==0.4.12 (not installed, no network) -- TEST INFRASTRUCTURE ONLY.

Only the symbols the reference imports are provided:
  * OCR/OmniParser/model/backbone/swin_transformer.py:14  -> DropPath, to_2tuple, trunc_normal_
  * OCR/MGP-STR/modules/mgp_str.py:19-21                  -> VisionTransformer, _cfg, register_model, create_model
Used by oracle/gen_golden.py to import the *unmodified* reference modules in this container.
"""
