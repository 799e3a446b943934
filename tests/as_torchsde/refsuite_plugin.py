"""pytest plugin loaded by tests/reference_suite.py: the reference's tests that take no `device` parameter build
their tensors on the default device — make that the GPU, since the package under test has no CPU path."""
import torch


def pytest_configure(config):
    if torch.cuda.is_available():
        torch.set_default_device('cuda')
