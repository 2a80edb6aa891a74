"""Second-stage (CORAL) test-set sweep — scripts/LTeval.py of the reference; see `eval.py` here for the flow."""
from .eval import main as _main


def main(argv=None):
    return _main(argv, second_stage=True)


if __name__ == "__main__":
    main()
