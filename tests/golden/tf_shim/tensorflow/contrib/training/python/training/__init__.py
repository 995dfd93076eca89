"""shim package."""
