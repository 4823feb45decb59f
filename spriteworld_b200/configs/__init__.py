"""Task configurations: modules exposing get_config(mode) -> kwargs of Environment."""
