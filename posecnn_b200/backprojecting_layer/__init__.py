# mirrors lib/backprojecting_layer/__init__.py of the reference
