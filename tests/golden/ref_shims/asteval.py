"""Stand-in for asteval: expressions are parsed with `ast` and evaluated with `eval` inside the
symbol table (adequate for the trusted expression strings of the reference's own configs)."""
import ast
import keyword
import re

import numpy as np

_NAME = re.compile(r'^[a-zA-Z_][a-zA-Z0-9_]*$')


def valid_symbol_name(name):
    return isinstance(name, str) and bool(_NAME.match(name)) and not keyword.iskeyword(name)


def get_ast_names(tree):
    return sorted({node.id for node in ast.walk(tree) if isinstance(node, ast.Name)})


class Interpreter(object):
    def __init__(self, *args, **kwargs):
        self.symtable = {name: getattr(np, name) for name in
                         ('sin', 'cos', 'tan', 'exp', 'log', 'log10', 'sqrt', 'arctan2', 'pi', 'abs', 'inf')}
        self._builtin = set(self.symtable)
        self.error = []
        self.error_msg = None

    def user_defined_symbols(self):
        return [k for k in self.symtable if k not in self._builtin]

    def parse(self, text):
        return ast.parse(text.strip(), mode='eval')

    def run(self, tree, **kwargs):
        return eval(compile(tree, '<expr>', 'eval'), {'__builtins__': {}}, self.symtable)

    def eval(self, text, **kwargs):
        return self.run(self.parse(text))

    def __call__(self, text, **kwargs):
        return self.eval(text)
