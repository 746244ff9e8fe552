# Test-infrastructure shim (NOT the product): lets the reference's modeling_nano.py import on a
# box without the mamba_ssm wheel so that its pure-PyTorch torch_forward can generate golden vectors.
