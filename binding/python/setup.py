"""Packaging of the Python binding (reference: binding/python/setup.py:10-30)."""
from setuptools import find_packages, setup

setup(
    name="multiverso-python",
    version="0.1.0",
    description="Python binding of multiverso-b200 (API compatible with Microsoft/multiverso's binding)",
    packages=find_packages(),
    install_requires=["numpy", "torch"],
)
