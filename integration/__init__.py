"""Reference-side binding of libnngpara.so (INTEGRATION.md section 3 as code)."""
