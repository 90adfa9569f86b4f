"""Shim of diffusers.pipelines (import resolution only)."""
