"""Drop-in `networks` package (`networks.classifiers`, `networks.losses`); anything else
(`networks.cpc`, `networks.apc`) resolves from a reference checkout later on sys.path."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
