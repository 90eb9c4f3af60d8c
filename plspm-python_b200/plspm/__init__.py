"""plspm -- drop-in host package of the B200-native PLS-PM engine.

Keeps the module paths and public names of GoogleCloudPlatform/plspm-python
(plspm.plspm.Plspm, plspm.config.{Config, Structure, MV}, plspm.scheme.Scheme,
plspm.mode.Mode, plspm.scale.Scale, plspm.bootstrap.Bootstrap, ...), and runs the
weight-estimation hot path on the GPU through libplspm_b200.so.
"""
name = "plspm"
