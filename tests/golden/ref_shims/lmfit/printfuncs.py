def params_html_table(params):
    return '<table></table>'
