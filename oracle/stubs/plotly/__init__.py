"""No-op plotly stand-in so the reference front-end imports without network deps (test scaffolding)."""
