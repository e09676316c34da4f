# mirrors lib/projecting_layer/__init__.py of the reference
