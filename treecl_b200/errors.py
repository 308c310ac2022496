"""Error types kept from the reference for this path (treeCl/errors.py:50-60,128-131)."""


class OptionError(Exception):
    def __init__(self, option, choices):
        self.value = option
        self.choices = choices

    def __str__(self):
        return '\'{0}\' is not a valid option. Choose from: {1}'.format(self.value, self.choices)


def optioncheck(option, choices):
    if option not in choices:
        raise OptionError(option, choices)
    return option
