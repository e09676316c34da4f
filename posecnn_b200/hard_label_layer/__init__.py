# mirrors lib/hard_label_layer/__init__.py of the reference
