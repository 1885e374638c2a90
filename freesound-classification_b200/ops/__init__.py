"""Drop-in `ops` package: same module names as the reference (`ops.utils`, `ops.audio`,
`ops.padding`, `ops.training`, `ops.transforms`).  Modules this package does not re-implement
(`ops.folds`, ...) resolve from a reference checkout placed LATER on sys.path."""
import pkgutil

__path__ = pkgutil.extend_path(__path__, __name__)
